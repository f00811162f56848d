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
Multi-GPU (one process per GPU, torch.distributed initialised - e.g. under torchrun): `main` shards the work and
all-reduces only the integer counters (`evaluate_sharded`): the wake-word clips go to the ranks in contiguous runs
(each rank starts with the window carry the clips before its run leave behind, computed from their lengths alone),
the long false-accept wav is cut into one time chunk per rank with a PCM halo of (L-1)*160+352 samples plus the
16 / 14 posteriors of the 30-tap smoothing, and the chunk counts add up exactly to the single-process counts.
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


def clip_num_samples(item, sample_rate: int) -> int:
    """Length of a clip without decoding it (wav header) - what the sharded evaluation needs of the clips of OTHER ranks."""
    if isinstance(item, np.ndarray):
        return int(item.shape[0])
    with wave.open(str(item), "rb") as f:
        return int(f.getnframes())


def load_wav_range(item, start: int, stop: int, sample_rate: int) -> np.ndarray:
    """Samples [start, stop) of a clip as float32 (zeros outside the clip): a rank's time chunk of the long FAR wav."""
    n = clip_num_samples(item, sample_rate)
    out = np.zeros((max(stop - start, 0),), np.float32)
    a, b = max(start, 0), min(stop, n)
    if b <= a:
        return out
    if isinstance(item, np.ndarray):
        out[a - start:b - start] = item[a:b].astype(np.float32, copy=False)
        return out
    with wave.open(str(item), "rb") as f:
        if f.getsampwidth() != 2 or f.getframerate() != sample_rate:
            raise ValueError("%s: only 16-bit PCM wavs at %d Hz are supported" % (item, sample_rate))
        f.setpos(a)
        raw = np.frombuffer(f.readframes(b - a), dtype=np.int16)
        if f.getnchannels() > 1:
            raw = raw.reshape(-1, f.getnchannels()).astype(np.float32).mean(axis=1)
    out[a - start:b - start] = (raw.astype(np.float32) / 32768.0).astype(np.float32)
    return out


def carry_before_clips(n_samples: Sequence[int], sample_rate: int, frame_length: int) -> List[int]:
    """Samples the shared 512-window still holds when clip i starts (evaluate_models.py:30: ONE Filter for all clips).
    Every padded clip ends in sample_rate // 2 >= 512 zeros, so the carried samples are zeros and only their NUMBER
    matters (it shifts the clip's frame grid): a function of the lengths of the clips before."""
    out, carry = [], 0
    for n in n_samples:
        out.append(carry)
        padded = -(-(int(n) + 2 * (sample_rate // 2)) // frame_length) * frame_length
        total = carry + padded
        nf = 0 if total < 512 else (total - 512) // HOP + 1
        carry = total - nf * HOP if nf else total
    return out


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
                  pre_emphasis: float = 0.0, carry_len: int = 0):
    """evaluate_models.py:26-108.  `test_files` may hold wav paths or float arrays.

    Returns a list: one maximum per clip for eval_type == "false_negatives", otherwise
    the concatenated posterior trajectories.  The reference creates ONE Filter for all
    clips (:30), so the 512-sample window still holds the tail of the previous clip when
    the next one starts; that carry is reproduced here (it shifts the frame grid of
    every clip after the first by the 480 zeros left in the window).  `carry_len`: zeros already in the window when the
    first clip starts (a rank's run of clips in the sharded evaluation, `carry_before_clips`)."""
    eng = engine or _cabi.engine_for_dir(models_dir, model_type)
    torch = eng.torch
    L = eng.L
    frame_length = sample_rate // 1000 * frame_width
    all_posterior: List = []
    carry = np.zeros((int(carry_len),), np.float32)
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


def far_chunk_posteriors(far_item, frame_width, sample_rate, rank: int, world: int, engine: "_cabi.Engine"):
    """Rank `rank`'s time chunk of the ONE long false-accept clip (evaluate_models.py:309-318).  The global trajectory has
    n_win posteriors (window j = mel frames [2j, 2j + L) of the padded stream); the rank counts windows [b, e) and
    computes [b - lo, e + hi) (dist.time_chunks: 16 / 14 posteriors of smoothing + edge halo), i.e. the PCM range
    [2 (b - lo) 160, (2 (e + hi - 1) + L - 1) 160 + 512) - chunk starts are multiples of 320 samples, so the chunk's
    window grid is the global one.  Returns (posteriors of the chunk, lo, hi, n_win)."""
    from . import dist
    frame_length = sample_rate // 1000 * frame_width
    half = sample_rate // 2
    n = clip_num_samples(far_item, sample_rate)
    n_pad = -(-(n + 2 * half) // frame_length) * frame_length
    n_win = engine.num_windows(engine.num_frames(n_pad), INFERENCE_HOP)
    b, e, lo, hi = dist.time_chunks(n_win, world)[rank]
    j0, j1 = b - lo, e + hi
    if j1 <= j0:
        return np.zeros((0,), np.float32), 0, 0, n_win
    p0 = INFERENCE_HOP * j0 * HOP
    p1 = (INFERENCE_HOP * (j1 - 1) + engine.L - 1) * HOP + 512
    chunk = load_wav_range(far_item, p0 - half, p1 - half, sample_rate)       # zeros where the padding is
    post = engine.pipeline(engine.torch.from_numpy(chunk[None]).to(engine.device), INFERENCE_HOP, 0.0).cpu().numpy()[0]
    assert post.shape[0] == j1 - j0
    return post, lo, hi, n_win


def evaluate_sharded(models_dir, model_type, wakeword_items, far_item, frame_width, sample_rate, thresholds,
                     rank: int = None, world: int = None, engine: "_cabi.Engine" = None, windowsize: int = 30,
                     reduce_over_ranks: bool = True):
    """The FRR / FAR numerators of evaluate_models.py:280-327 for this rank's share: wake-word clips
    [shard_range(rank)] with the right window carry, its time chunk of the long FAR clip; ONE all-reduce of the two
    int64 counter vectors (NCCL on GPUs).  Returns (accepts_pos[t], edges_neg[t]) - identical on every rank, and
    identical to the single-process sweep."""
    from . import dist
    if rank is None or world is None:
        rank, world = dist.rank_world()
    eng = engine or _cabi.engine_for_dir(models_dir, model_type)
    frame_length = sample_rate // 1000 * frame_width
    items = list(wakeword_items)
    carries = carry_before_clips([clip_num_samples(it, sample_rate) for it in items], sample_rate, frame_length)
    b, e = dist.shard_range(len(items), rank, world)
    pos = get_posterior(models_dir, model_type, "false_negatives", items[b:e], frame_width, sample_rate, engine=eng,
                        carry_len=carries[b] if b < len(items) else 0) if e > b else []
    kp = np.asarray(pos, np.float32)
    torch = eng.torch
    if kp.size:
        acc = eng.eval_counts(kp, np.arange(kp.size + 1), thresholds, "frr_max")
    else:
        acc = torch.zeros((len(thresholds),), dtype=torch.int64, device=eng.device)
    part, lo, hi, _ = far_chunk_posteriors(far_item, frame_width, sample_rate, rank, world, eng)
    if part.size - lo - hi > 0:
        edg = eng.eval_counts(part, [0, part.size], thresholds, "far_edges", windowsize, halo_lo=[lo], halo_hi=[hi])
    else:
        edg = torch.zeros_like(acc)
    if reduce_over_ranks:
        acc, edg = dist.all_reduce_counters(acc, edg)
    return acc.cpu().numpy(), edg.cpu().numpy()


def plot_FRR_FAR(keyword_posteriors, no_keyword_posteriors, num_wakewords, total_duration_hrs, model_type,
                 engine: "_cabi.Engine" = None, show: bool = True, counts=None):
    """evaluate_models.py:183-252.  Returns (thresholds, FRR, FAR) — the reference returns
    nothing and only plots; the plots are drawn when matplotlib is importable."""
    thresholds = np.arange(0.5, 0.99999, 0.005)
    if engine is None and counts is None:
        engs = [e for e in _cabi._ENGINES.values() if e.ctx is not None]
        if not engs:
            raise RuntimeError("plot_FRR_FAR needs an Engine (run get_posterior first or pass engine=)")
        engine = engs[0]
    print('Sweeping thresholds over posteriors')
    acc, edg = counts if counts is not None else sweep_counts(keyword_posteriors, no_keyword_posteriors, thresholds, engine)
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
    from . import dist
    rank, world = dist.rank_world()
    if world > 1:
        # one process per GPU: shard, count locally, all-reduce the counters; every rank gets the same curves, rank 0 plots
        eng = _cabi.engine_for_dir(args.models_dir, args.model_type, device=int(os.environ.get("LOCAL_RANK", "0")))
        counts = evaluate_sharded(args.models_dir, args.model_type, wakeword_paths, str(FAR_path), args.frame_width,
                                  args.sample_rate, np.arange(0.5, 0.99999, 0.005), rank, world, eng)
        plot_FRR_FAR(None, None, num_wakewords, total_duration_hrs, args.model_type, show=rank == 0, counts=counts)
        return 0
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
