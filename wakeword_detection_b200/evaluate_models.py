'''
FAR/FRR evaluation entry point — drop-in for the reference's utils/evaluate_models.py
(reference lines cited per function).  Same functions, arguments and CLI flags:

    python -m wakeword_detection_b200.evaluate_models --model_type CRNN \\
        --models_dir <dir with filter/encode/detect .tflite or weights.npz> --data_dir <hey-snips dir>

The posteriors come from the CUDA pipeline (filter -> encode -> detect over all
windows of a batch of clips at once) instead of one TFLite invoke per window; the
threshold sweep runs on the device.  librosa / pydub / matplotlib are optional here:
wavs are read with the standard `wave` module (16-bit PCM at the requested rate, scaled
by 1/32768 like librosa does), plotting is skipped when matplotlib is missing.
With torch.distributed initialised, clips are sharded over ranks and only the integer
counters are all-reduced (dist.py).
'''
from __future__ import annotations

import argparse
import json
import logging
import os
import pickle
import sys
import wave
from pathlib import Path
from typing import List, Sequence

import numpy as np

from . import _cabi, weights as W

logging.basicConfig(level=logging.INFO)

HOP = 160
INFERENCE_HOP = 2          # evaluate_models.py:42


def load_wav(path, sample_rate: int) -> np.ndarray:
    """librosa.load(path, sr=sample_rate) for 16-bit PCM wavs already at that rate."""
    with wave.open(str(path), "rb") as f:
        if f.getsampwidth() != 2:
            raise ValueError("%s: only 16-bit PCM wavs are supported" % path)
        if f.getframerate() != sample_rate:
            raise ValueError("%s: sample rate %d != %d (resampling needs librosa)"
                             % (path, f.getframerate(), sample_rate))
        raw = np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16)
        if f.getnchannels() > 1:
            raw = raw.reshape(-1, f.getnchannels()).astype(np.float32).mean(axis=1) / 32768.0
            return raw.astype(np.float32)
    return (raw.astype(np.float32) / 32768.0).astype(np.float32)


def _clip_samples(item, sample_rate: int) -> np.ndarray:
    if isinstance(item, np.ndarray):
        return item.astype(np.float32, copy=False)
    return load_wav(item, sample_rate)


def _padded_stream(samples: np.ndarray, sample_rate: int, frame_length: int) -> np.ndarray:
    """evaluate_models.py:52-61: sr//2 zeros on both sides, then whole frame_length
    chunks (the last one zero-padded)."""
    n = samples.shape[0] + 2 * (sample_rate // 2)
    n_pad = -(-n // frame_length) * frame_length
    out = np.zeros((n_pad,), np.float32)
    out[sample_rate // 2: sample_rate // 2 + samples.shape[0]] = samples
    return out


def get_posterior(models_dir, model_type, eval_type, test_files, frame_width, sample_rate,
                  examine_audio=False, engine: "_cabi.Engine" = None, batch_clips: int = 512,
                  pre_emphasis: float = 0.0):
    """evaluate_models.py:26-108.  `test_files` may hold wav paths or float arrays.

    Returns a list: one maximum per clip for eval_type == "false_negatives", otherwise
    the concatenated posterior trajectories.  The reference creates ONE Filter for all
    clips (:30), so the 512-sample window still holds the tail of the previous clip when
    the next one starts; that carry is reproduced here (it shifts the frame grid of
    every clip after the first by the 480 zeros left in the window)."""
    eng = engine or _cabi.engine_for_dir(models_dir, model_type)
    torch = eng.torch
    L = eng.L
    frame_length = sample_rate // 1000 * frame_width
    all_posterior: List = []
    carry = np.zeros((0,), np.float32)
    prev = 0.0
    files = list(test_files)
    for b0 in range(0, len(files), batch_clips):
        streams = []
        for item in files[b0:b0 + batch_clips]:
            x = _padded_stream(_clip_samples(item, sample_rate), sample_rate, frame_length)
            if pre_emphasis != 0.0:
                y = (x - np.float32(pre_emphasis) * np.concatenate([[np.float32(prev)], x[:-1]])).astype(np.float32)
            else:
                y = x
            prev = float(x[-1])
            s = np.concatenate([carry, y])
            nf = eng.num_frames(s.shape[0])
            carry = s[nf * HOP:] if nf else s
            streams.append((s, nf))
        n_max = max(s.shape[0] for s, _ in streams)
        host = np.zeros((len(streams), n_max), np.float32)
        for i, (s, _) in enumerate(streams):
            host[i, :s.shape[0]] = s
        post = eng.pipeline(torch.from_numpy(host).to(eng.device), INFERENCE_HOP, 0.0).cpu().numpy()
        for i, (s, nf) in enumerate(streams):
            nw = eng.num_windows(nf, INFERENCE_HOP)
            p = post[i, :nw]
            if eval_type == "false_negatives":
                all_posterior.append(np.max(p))       # raises on an empty clip like the reference (:99)
            else:
                all_posterior.extend(list(p))
    return all_posterior


def testset_files(base_path):
    """evaluate_models.py:137-145."""
    json_path = base_path + "test.json"
    test_data = json.load(open(json_path, 'r'))
    wakeword_files = [base_path + path["audio_file_path"] for path in test_data if path["is_hotword"]]
    not_wakeword_files = [base_path + path["audio_file_path"] for path in test_data if not path["is_hotword"]]
    return wakeword_files, not_wakeword_files


def concatenate_FA(wavs: Sequence[np.ndarray], num_files, FAR_path, sample_rate: int = 16000):
    """evaluate_models.py:148-158: the first num_files negatives joined by 100 ms of
    silence, written as one 16-bit wav.  `wavs` are int16 arrays here (no pydub)."""
    gap = np.zeros((sample_rate // 10,), np.int16)
    parts = [np.asarray(wavs[0], np.int16)]
    for w in wavs[1:num_files]:
        parts += [gap, np.asarray(w, np.int16)]
    data = np.concatenate(parts)
    with wave.open(str(FAR_path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sample_rate)
        f.writeframes(data.tobytes())


def load_posteriors(models_dir, model_type, frame_width, sample_rate, eval_type, input_path, out_path,
                    examine_audio=False):
    """evaluate_models.py:161-174 (same pickle cache files)."""
    out_path = Path(out_path)
    if out_path.exists():
        with open(out_path, 'rb') as f:
            posteriors = pickle.load(f)
    else:
        posteriors = get_posterior(models_dir, model_type, eval_type, input_path, frame_width, sample_rate,
                                   examine_audio)
        with open(out_path, 'wb') as f:
            pickle.dump(posteriors, f)
    return np.squeeze(np.array(posteriors))


def duration_test(FAR_path, sample_rate):
    """evaluate_models.py:177-180."""
    with wave.open(str(FAR_path), "rb") as f:
        return f.getnframes() / sample_rate


def sweep_counts(keyword_posteriors, no_keyword_posteriors, thresholds, engine: "_cabi.Engine",
                 windowsize: int = 30, reduce_over_ranks: bool = False):
    """Integer numerators of the sweep on the device: (accepts_pos[t], edges_neg[t])."""
    kp = np.atleast_1d(np.asarray(keyword_posteriors, np.float32))
    nk = np.atleast_1d(np.asarray(no_keyword_posteriors, np.float32))
    acc = engine.eval_counts(kp, np.arange(kp.size + 1), thresholds, "frr_max")
    if nk.size:
        edg = engine.eval_counts(nk, [0, nk.size], thresholds, "far_edges", windowsize)
    else:
        edg = engine.torch.zeros_like(acc)
    if reduce_over_ranks:
        from . import dist
        acc, edg = dist.all_reduce_counters(acc, edg)
    return acc.cpu().numpy(), edg.cpu().numpy()


def plot_FRR_FAR(keyword_posteriors, no_keyword_posteriors, num_wakewords, total_duration_hrs, model_type,
                 engine: "_cabi.Engine" = None, show: bool = True):
    """evaluate_models.py:183-252.  Returns (thresholds, FRR, FAR) — the reference returns
    nothing and only plots; the plots are drawn when matplotlib is importable."""
    thresholds = np.arange(0.5, 0.99999, 0.005)
    if engine is None:
        engs = [e for e in _cabi._ENGINES.values() if e.ctx is not None]
        if not engs:
            raise RuntimeError("plot_FRR_FAR needs an Engine (run get_posterior first or pass engine=)")
        engine = engs[0]
    print('Sweeping thresholds over posteriors')
    acc, edg = sweep_counts(keyword_posteriors, no_keyword_posteriors, thresholds, engine)
    FRR = [(num_wakewords - int(a)) / num_wakewords for a in acc]
    FAR = [int(e) / total_duration_hrs for e in edg]
    if show:
        try:
            from matplotlib import pyplot as plt  # type: ignore
        except Exception:
            plt = None
        if plt is not None:   # pragma: no cover - matplotlib is not installed in this image
            for x, y, xl, yl in ((thresholds, FRR, "Posterior Threshold", "False Rejection Rate"),
                                 (thresholds, FAR, "Posterior Threshold", "False Accepts per Hour"),
                                 (FAR, FRR, "False Alarms per Hour", "False Rejection Rate")):
                fig, ax = plt.subplots(1, 1)
                ax.set_facecolor('lightgray')
                plt.plot(x, y, label=model_type)
                plt.ylabel(yl)
                plt.xlabel(xl)
                plt.grid(color='white')
                plt.legend()
                plt.tight_layout()
                plt.show()
                plt.close()
    return thresholds, FRR, FAR


def parse_args(argv=None):
    """evaluate_models.py:256-278 (same flags and defaults, except a portable models_dir)."""
    parser = argparse.ArgumentParser(description='Evaluates wakeword model(s), reports useful metrics.')
    parser.add_argument('--model_type', type=str, default='CRNN', choices=['CRNN', 'Wavenet'],
                        help='Model type being evaluated.')
    parser.add_argument('--models_dir', type=str,
                        default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             'weights', 'CRNN_arik_original') + os.sep,
                        help='Directory where trained models are stored.')
    parser.add_argument('--data_dir', type=str, default='data/hey_snips_research_6k_en_train_eval_clean_ter/',
                        help='Directory with Hey Snips raw dataset')
    parser.add_argument('--eval_dir', type=str, default='data/evaluation/',
                        help='Directory to save and load concatenated wav files from')
    parser.add_argument('--pos_samples', type=str, default='hey_snips_long.wav',
                        help='File for concatenated positive class samples')
    parser.add_argument('--neg_samples', type=str, default='not_hey_snips_long.wav',
                        help='File for concatenated negative class samples')
    parser.add_argument('--sample_rate', type=int, default=16000, help='Sample rate for audio (Hz)')
    parser.add_argument('--frame_width', type=int, default=20, help='Frame width for audio in (ms)')
    parser.add_argument('--examine_audio', default=False, action='store_true',
                        help='Flag to examine problematic audio clips')
    args = parser.parse_args(argv)
    assert Path(args.models_dir).exists(), "Directory for TF-Lite models and results is not found!"
    return args


def main(args) -> int:
    """evaluate_models.py:280-327."""
    FAR_path = Path(os.path.join(args.eval_dir, args.neg_samples))
    wakeword_paths, not_wakeword_paths = testset_files(args.data_dir)
    num_wakewords = len(wakeword_paths)
    if not FAR_path.exists():
        try:
            os.mkdir(args.eval_dir)
        except FileExistsError:
            print("directory for eval files exists")
        print('Loading all files for FA test set')
        wavs = []
        for p in not_wakeword_paths[:num_wakewords]:
            with wave.open(p, "rb") as f:
                wavs.append(np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16))
        concatenate_FA(wavs, num_wakewords, FAR_path, args.sample_rate)
        del wavs
    print('Calculating total duration of FA test set')
    total_duration_hrs = duration_test(FAR_path, args.sample_rate) / 3600
    print(f'Total duration of FA set is {total_duration_hrs:.2f} hrs')
    pos = load_posteriors(args.models_dir, args.model_type, args.frame_width, args.sample_rate,
                          "false_negatives", wakeword_paths,
                          Path(os.path.join(args.models_dir, args.model_type + "_all_wakeword.pkl")),
                          args.examine_audio)
    neg = load_posteriors(args.models_dir, args.model_type, args.frame_width, args.sample_rate,
                          "false_accepts", [str(FAR_path)],
                          Path(os.path.join(args.models_dir, args.model_type + "_no_wakeword.pkl")),
                          args.examine_audio)
    eng = _cabi.engine_for_dir(args.models_dir, args.model_type)
    plot_FRR_FAR(pos, neg, num_wakewords, total_duration_hrs, args.model_type, engine=eng)
    return 0


if __name__ == '__main__':
    sys.exit(main(parse_args()))
