"""Generates int16 PCM clips that make each shipped model fire (posterior near 1),
by gradient ascent through a torch-CPU copy of the pipeline.  Synthetic noise never
triggers either model (posteriors ~0), so without these clips the threshold /
trigger / FAR / FRR logic would only ever be tested on zeros (SURVEY.md §7 risks).

Run here (needs /root/reference for the trained weights):
    python tests/golden/make_adversarial.py
Writes tests/golden/wake_{crnn,wavenet}_pcm.npy (int16) and wake_{crnn,wavenet}_mel.npy
(a handful of mel windows with posteriors spread over [0,1]).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from wakeword_detection_b200 import weights as W, synth  # noqa: E402
from oracle import restated as R  # noqa: E402

REF = "/root/reference/tf_lite_models/"


def T(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32)


def mel_torch(x, w):
    frames = x.unfold(0, 512, 160)
    win = torch.tensor(np.hanning(512), dtype=torch.float32)
    mag = torch.fft.rfft(frames * win, n=512).abs()
    y = mag @ T(w["mel_w"]).T
    y = torch.clamp(y, min=float(w["mel_floor"]))
    return (torch.log(y) - float(w["mel_log_offset"])) * float(w["mel_scale"])


def gru_torch(seq, W_, U, bi, br, reverse):
    B, Tn, _ = seq.shape
    xw = seq @ T(W_).T + T(bi)
    h = torch.zeros((B, 32))
    outs = [None] * Tn
    for t in (range(Tn - 1, -1, -1) if reverse else range(Tn)):
        hu = h @ T(U).T + T(br)
        z = torch.sigmoid(xw[:, t, :32] + hu[:, :32])
        r = torch.sigmoid(xw[:, t, 32:64] + hu[:, 32:64])
        c = torch.tanh(xw[:, t, 64:] + r * hu[:, 64:])
        h = z * h + (1 - z) * c
        outs[t] = h
    return torch.stack(outs, 1)


def crnn_logit(mel, w):
    x = mel.transpose(1, 2)[:, None]                       # [B,1,40,151]
    x = torch.nn.functional.pad(x, (6, 7, 1, 2))
    y = torch.nn.functional.conv2d(x, T(w["conv_w"])[:, None], T(w["conv_b"]), stride=(2, 8))
    y = torch.relu(y).permute(0, 3, 2, 1).reshape(mel.shape[0], 19, 640)
    f1 = gru_torch(y, w["gru1_f_w"], w["gru1_f_u"], w["gru1_f_bi"], w["gru1_f_br"], False)
    b1 = gru_torch(y, w["gru1_b_w"], w["gru1_b_u"], w["gru1_b_bi"], w["gru1_b_br"], True)
    s1 = torch.cat([f1, b1], 2)
    f2 = gru_torch(s1, w["gru2_f_w"], w["gru2_f_u"], w["gru2_f_bi"], w["gru2_f_br"], False)
    b2 = gru_torch(s1, w["gru2_b_w"], w["gru2_b_u"], w["gru2_b_bi"], w["gru2_b_br"], True)
    e = torch.cat([f2[:, -1], b2[:, 0]], 1)
    h = torch.relu(e @ T(w["det1_w"]).T + T(w["det1_b"]))
    z = h @ T(w["det2_w"]).T + T(w["det2_b"])
    return z[:, 0] if z.shape[1] == 1 else z[:, 1] - z[:, 0]


def wavenet_logit(mel, w):
    x = torch.relu(mel @ T(w["in_w"]).T + T(w["in_b"]))
    Tn = mel.shape[1]
    out = 0
    for k in range(24):
        d = int(w["dilation"][k])
        u = x * T(w["bn_mul"][k]) + T(w["bn_add"][k])
        up = torch.nn.functional.pad(u, (0, 0, 2 * d, 0))
        taps = torch.cat([up[:, j * d:j * d + Tn] for j in range(3)], 2)
        g = torch.tanh(taps @ T(w["tanh_w"][k].reshape(16, 48)).T + T(w["tanh_b"][k])) * \
            torch.sigmoid(taps @ T(w["sig_w"][k].reshape(16, 48)).T + T(w["sig_b"][k]))
        out = out + torch.relu(g @ T(w["skip_w"][k]).T + T(w["skip_b"][k]))
        if k < 23:
            x = torch.relu(g @ T(w["res_w"][k]).T + T(w["res_b"][k])) + x
    h = torch.relu(torch.relu(out) @ T(w["det1_w"]).T + T(w["det1_b"]))
    z = (h @ T(w["det2_w"]).T + T(w["det2_b"])).max(1).values
    return z[:, 1] - z[:, 0]


def main():
    torch.manual_seed(0)
    for name, sub, typ in (("crnn", "CRNN", "CRNN"), ("wavenet", "Wavenet", "Wavenet")):
        w = W.load_model_dir(REF + sub, typ)
        L = int(w["mel_length"])
        logit_fn = crnn_logit if typ == "CRNN" else wavenet_logit
        n = 16000 * 2 + 3200
        x0 = synth.stream_float(n, 2, seed=7, stream=1) * 0.5
        x = torch.tensor(x0, dtype=torch.float32, requires_grad=True)
        opt = torch.optim.Adam([x], lr=2e-3)
        for it in range(400):
            opt.zero_grad()
            xp = torch.nn.functional.pad(torch.clamp(x, -1, 1), (8000, 8000))
            mel = mel_torch(xp, w)
            j0 = (mel.shape[0] - L) // 2 // 2 * 2
            wins = torch.stack([mel[j:j + L] for j in range(j0 - 4, j0 + 6, 2)])
            lg = logit_fn(wins, w)
            loss = -torch.clamp(lg, max=6.0).mean()
            loss.backward()
            opt.step()
            if it % 50 == 0:
                print(name, it, lg.detach().numpy().round(2))
        pcm = np.clip(np.rint(x.detach().numpy().astype(np.float64) * 32767), -32768, 32767).astype(np.int16)
        np.save(os.path.join(HERE, "wake_%s_pcm.npy" % name), pcm)
        # report what the numpy oracle sees on the quantised clip, eval framing
        post = np.array(R.get_posterior([pcm.astype(np.float32) / 32768.0], w, "false_accepts"))
        print(name, "oracle posteriors on int16 clip: max %.4f, >0.5: %d of %d" %
              (post.max(), (post > 0.5).sum(), post.size))
        # mel windows with posteriors spread over [0, 1]: blend the best window with silence
        xp = np.pad(pcm.astype(np.float32) / 32768.0, (8000, 8000))
        mel = R.mel_stream(xp, w)
        jbest = int(np.argmax(post)) * 2
        best = mel[jbest:jbest + L]
        rng = np.random.default_rng(3)
        noise = R.mel_stream(synth.stream_float(16000 * 3, 0, 3, 0).astype(np.float32), w)[:L]
        wins = np.stack([best * a + noise * (1 - a) for a in np.linspace(0, 1, 24)] +
                        [mel[j:j + L] for j in range(0, mel.shape[0] - L, 8)][:24]).astype(np.float32)
        pw = R.posterior(wins, w)
        print(name, "blend posteriors", pw.round(3))
        np.save(os.path.join(HERE, "wake_%s_mel.npy" % name), wins)


if __name__ == "__main__":
    main()
