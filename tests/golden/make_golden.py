"""Golden vectors from the reference's OWN Python glue, run in this container.

The reference has no tests or fixtures.  Its Python glue — ring
buffers, pre-emphasis, per-sample framing, window bookkeeping, trigger logic,
FAR/FRR sweeps — *is* importable here once the third-party pieces are stubbed:

  tflite_runtime.interpreter.Interpreter -> oracle/tflite_literal.py (literal
      execution of the shipped .tflite graphs; TFLite itself is not installed)
  librosa.load            -> returns in-memory float clips keyed by "path"
  pydub / matplotlib      -> inert stubs (playback, plotting)

So the vectors below pin the *glue semantics* against unmodified reference code
(/root/reference/spokestack/wakeword/tflite.py, utils/tf_lite/filter.py,
utils/evaluate_models.py, utils/plot_eval_models.py); the op arithmetic stays
pinned only by the literal interpreter (see oracle/tflite_literal.py header).

    python tests/golden/make_golden.py      # writes tests/golden/reference_glue.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle.tflite_literal import LiteralInterpreter  # noqa: E402
from wakeword_detection_b200 import synth  # noqa: E402

CLIPS = {}


def install_stubs():
    class Interpreter:
        def __init__(self, model_path, **kw):
            self._m = LiteralInterpreter(model_path)
            self._in = {}
            self._out = []

        def get_input_details(self):
            return self._m.input_details

        def get_output_details(self):
            return self._m.output_details

        def allocate_tensors(self):
            pass

        def set_tensor(self, index, value):
            self._in[index] = np.array(value)

        def invoke(self):
            args = [self._in[d["index"]] for d in self._m.input_details]
            self._out = self._m(*args)

        def get_tensor(self, index):
            for d, o in zip(self._m.output_details, self._out):
                if d["index"] == index:
                    return np.array(o)
            raise KeyError(index)

    rt = types.ModuleType("tflite_runtime")
    rti = types.ModuleType("tflite_runtime.interpreter")
    rti.Interpreter = Interpreter
    rt.interpreter = rti
    sys.modules["tflite_runtime"] = rt
    sys.modules["tflite_runtime.interpreter"] = rti

    pydub = types.ModuleType("pydub")

    class AudioSegment:
        @staticmethod
        def from_mp3(p):
            return p

        @staticmethod
        def from_wav(p):
            return p

        @staticmethod
        def silent(duration=0):
            return None

    pydub.AudioSegment = AudioSegment
    pb = types.ModuleType("pydub.playback")
    pb.play = lambda seg: None
    pydub.playback = pb
    sys.modules["pydub"] = pydub
    sys.modules["pydub.playback"] = pb

    librosa = types.ModuleType("librosa")
    librosa.load = lambda path, sr=16000: (CLIPS[path].astype(np.float32).copy(), sr)
    sys.modules["librosa"] = librosa

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.captured = []

    class _Ax:
        def __getattr__(self, name):
            return lambda *a, **k: None

    def subplots(*a, **k):
        n = (a[0] if a else 1) * (a[1] if len(a) > 1 else 1)
        return None, (_Ax() if n == 1 else [_Ax() for _ in range(n)])

    plt.subplots = subplots
    plt.plot = lambda x, y, *a, **k: plt.captured.append((np.array(x), np.array(y)))
    for fn in ("ylabel", "xlabel", "grid", "legend", "tight_layout", "show", "close", "savefig",
               "xlim", "ylim"):
        setattr(plt, fn, lambda *a, **k: None)
    plt.axis = lambda *a, **k: (0, 1, 0, 1)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    return plt


def main():
    plt = install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "utils"))
    os.chdir(REF)                       # WakewordTrigger lists ./audio_responses
    from spokestack.wakeword.tflite import WakewordTrigger
    from spokestack.context import SpeechContext
    from tf_lite.filter import Filter
    import evaluate_models as EM
    import plot_eval_models as PM

    out = {}
    wake = {k: np.load(os.path.join(HERE, "wake_%s_pcm.npy" % k)) for k in ("crnn", "wavenet")}

    # ---- 1. Filter.filter_frame over ragged chunks, two pre-emphasis settings ----------
    x = np.concatenate([synth.stream_float(5000, c, 11, c) for c in (2, 0, 3, 4)]).astype(np.float32)
    x = np.clip(x, -1, 1)
    out["filter_in"] = x
    for tag, a in (("pe0", 0.0), ("pe97", 0.97)):
        f = Filter(pre_emphasis=a, model_dir=os.path.join(REF, "tf_lite_models/CRNN"))
        mels, counts, pos = [], [], 0
        for n in [320, 320, 7, 1000, 1, 512, 160, 159, 3000] + [320] * 40:
            chunk = x[pos:pos + n].copy()
            pos += n
            if chunk.size == 0:
                break
            got = f.filter_frame(chunk)
            counts.append(len(got))
            mels += [np.array(g) for g in got]
        out["filter_mel_" + tag] = np.stack(mels).astype(np.float32)
        out["filter_counts_" + tag] = np.array(counts, np.int32)
        out["filter_consumed_" + tag] = np.array(pos, np.int64)

    # ---- 2. WakewordTrigger streaming (hop 1, zero-prefilled frame window) ----------------
    for name, sub, typ in (("crnn", "CRNN", "CRNN"), ("wavenet", "Wavenet", "Wavenet")):
        trig = WakewordTrigger(model_dir=os.path.join(REF, "tf_lite_models", sub), model_type=typ)
        posts = []
        orig = trig.detect_model

        class Rec:
            input_details = orig.input_details
            output_details = orig.output_details

            def __call__(self, *a):
                r = orig(*a)
                posts.append(float(r[0][0][-1]))
                return r

        trig.detect_model = Rec()
        ctx = SpeechContext()
        pcm = np.concatenate([synth.stream_int16(3200, 2, 5, 0), wake[name][:16000 * 2]])
        n_frames = pcm.shape[0] // 320
        speech = np.ones(n_frames, bool)
        speech[:3] = False                       # VAD gate closed for the first 3 frames
        active_at = -1
        for i in range(n_frames):
            ctx.is_speech = bool(speech[i])
            trig(ctx, pcm[i * 320:(i + 1) * 320])
            if ctx.is_active and active_at < 0:
                active_at = i
                break                            # reference stops sampling once active
        out["trig_%s_pcm" % name] = pcm
        out["trig_%s_speech" % name] = speech
        out["trig_%s_post" % name] = np.array(posts, np.float32)
        out["trig_%s_active_at" % name] = np.array(active_at, np.int32)
        out["trig_%s_post_max" % name] = np.array(trig._posterior_max, np.float32)

    # ---- 3. get_posterior over three clips, both eval types -----------------------------------
    EM.tqdm = lambda it, **k: it
    for name, sub, typ in (("crnn", "wwdetect/CRNN/models/Arik_CRNN_data_original", "CRNN"),
                           ("wavenet", "tf_lite_models/Wavenet", "Wavenet")):
        mdir = os.path.join(REF, sub) + "/"
        if not os.path.exists(mdir + "filter.tflite"):
            import tempfile
            tmp = tempfile.mkdtemp()
            for f in ("encode.tflite", "detect.tflite"):
                os.symlink(mdir + f, os.path.join(tmp, f))
            os.symlink(os.path.join(REF, "tf_lite_models/CRNN/filter.tflite"), os.path.join(tmp, "filter.tflite"))
            mdir = tmp + "/"
        clips = [wake[name].astype(np.float32) / 32768.0,
                 np.clip(synth.stream_float(16000 + 4321, 2, 9, 1), -1, 1).astype(np.float32),
                 np.clip(synth.stream_float(16000 * 2 + 77, 0, 9, 2), -1, 1).astype(np.float32)]
        CLIPS.clear()
        for i, c in enumerate(clips):
            CLIPS["clip%d" % i] = c
            out["gp_%s_clip%d" % (name, i)] = c
        fn = EM.get_posterior(mdir, typ, "false_negatives", list(CLIPS), 20, 16000)
        fa = EM.get_posterior(mdir, typ, "false_accepts", list(CLIPS), 20, 16000)
        out["gp_%s_frr_max" % name] = np.array(fn, np.float32)
        out["gp_%s_far_traj" % name] = np.array(fa, np.float32)

        # ---- 4. plot_FRR_FAR numerators on those posteriors ------------------------------
        plt.captured.clear()
        EM.plot_FRR_FAR(np.array(fn), np.array(fa), len(fn), 0.5, typ)
        thr, frr = plt.captured[0]
        _, far = plt.captured[1]
        out["sweep_%s_thr" % name] = thr
        out["sweep_%s_frr" % name] = frr
        out["sweep_%s_far" % name] = far

    # ---- 5. plot_eval_models.threshold_accepts on a synthetic trajectory -------------------------
    rng = np.random.default_rng(5)
    traj = np.clip(np.convolve(rng.random(5000) ** 6, np.ones(9) / 3, "same"), 0, 1)
    thr = np.arange(0.5, 0.9905, 0.001)
    out["pm_traj"] = traj
    out["pm_thr"] = thr
    out["pm_accepts"] = np.array([PM.threshold_accepts(traj, t) for t in thr], np.int64)

    np.savez_compressed(os.path.join(HERE, "reference_glue.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
