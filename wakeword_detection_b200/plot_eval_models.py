"""Drop-in for the reference's multi-model result plots (reference: utils/plot_eval_models.py:70-129): `load_model_preds`
reads the pickles `evaluate_models.py` caches, `threshold_accepts` is the rising-edge counter, `process_results` sweeps
the 491 thresholds `arange(0.5, 0.9905, 0.001)` - here both numerators of every threshold come from ONE pass of the
counter kernels (csrc/counts.cu) instead of a Python loop over thresholds x posteriors.  Plotting (`plot_results`) runs
when matplotlib is importable."""
from __future__ import annotations

import os
import pickle

import numpy as np

from . import _cabi


def load_model_preds(model_types, results_dir, smooth_window=30):
    """:70-83."""
    assert len(model_types) == len(results_dir), 'Size mismatch between number of models and results!'
    results = {model: {} for model in model_types}
    for model, result_dir in zip(model_types, results_dir):
        results[model]['wakeword'] = pickle.load(open(os.path.join(result_dir, f'{model}_all_wakeword.pkl'), 'rb'))
        results[model]['not_wakeword'] = pickle.load(open(os.path.join(result_dir, f'{model}_no_wakeword.pkl'), 'rb'))
        results[model]['smooth_not_wakeword'] = np.convolve(results[model]['not_wakeword'],
                                                            np.ones((smooth_window,)) / smooth_window, mode='same')
    return results


def threshold_accepts(posteriors, threshold, engine: "_cabi.Engine" = None):
    """:85-97: rising edges of (posterior > threshold) along an (already smoothed) trajectory."""
    p = np.asarray(posteriors, np.float64)
    if engine is None:
        above = p > threshold
        return int(np.count_nonzero(above & ~np.concatenate([[False], above[:-1]])))
    return int(engine.eval_counts(p.astype(np.float32), [0, p.size], [threshold], "far_edges", smooth=1)[0])


def process_results(results, num_wakewords, not_wakeword_duration_hrs, engine: "_cabi.Engine" = None):
    """:99-129.  The false-accept count of a threshold is taken on the float64 smoothed trajectory like the reference
    does; the device kernel smooths the raw float32 trajectory itself in float64 (same 30-tap 'same' mean), so
    'not_wakeword' is what it reads when present."""
    if engine is None:
        engs = [e for e in _cabi._ENGINES.values() if e.ctx is not None]
        if not engs:
            raise RuntimeError("process_results needs an Engine (pass engine=)")
        engine = engs[0]
    for model in results:
        thresholds = np.arange(0.5, 0.9905, 0.001)
        print(f'Sweeping thresholds over posteriors for {model} model')
        kp = np.atleast_1d(np.asarray(results[model]['wakeword'], np.float32))
        acc = engine.eval_counts(kp, np.arange(kp.size + 1), thresholds, "frr_max").cpu().numpy()
        if 'not_wakeword' in results[model]:
            nk = np.atleast_1d(np.asarray(results[model]['not_wakeword'], np.float32))
            edg = engine.eval_counts(nk, [0, nk.size], thresholds, "far_edges", 30).cpu().numpy()
        else:
            sm = np.atleast_1d(np.asarray(results[model]['smooth_not_wakeword'], np.float32))
            edg = engine.eval_counts(sm, [0, sm.size], thresholds, "far_edges", 1).cpu().numpy()
        FRR = [(num_wakewords - int(a)) / num_wakewords for a in acc]
        FAR = [int(e) / not_wakeword_duration_hrs for e in edg]
        results[model]['FRR'] = sorted(FRR)[::-1]
        results[model]['FAR'] = sorted(FAR)
        results[model]['smooth_FAR'] = np.convolve(results[model]['FAR'], np.ones((30,)) / 30, mode='valid')
        results[model]['smooth_FRR'] = np.convolve(results[model]['FRR'], np.ones((30,)) / 30, mode='valid')
        results[model]['thresholds'] = thresholds
    return results


def plot_results(results):   # pragma: no cover - matplotlib is not installed in this image
    """:17-68 (skipped when matplotlib is missing)."""
    try:
        from matplotlib import pyplot as plt  # type: ignore
    except Exception:
        return
    for key, yl in (("FRR", "False Rejection Rate"), ("FAR", "False Accepts per Hour")):
        fig, ax = plt.subplots(1, 1)
        ax.set_facecolor('lightgray')
        for model in results:
            plt.plot(results[model]['thresholds'], results[model][key], label=model)
        plt.ylabel(yl)
        plt.xlabel("Posterior Threshold")
        plt.grid(color='white')
        plt.legend()
        plt.tight_layout()
        plt.show()
        plt.close()
    fig, ax = plt.subplots(1, 1)
    ax.set_facecolor('lightgray')
    for model in results:
        plt.plot(results[model]['smooth_FAR'], results[model]['smooth_FRR'], label=model)
    plt.xlabel("False Alarms per Hour")
    plt.ylabel("False Rejection Rate")
    plt.xlim(0, 12)
    plt.grid(color='white')
    plt.legend()
    plt.tight_layout()
    plt.savefig('far_frr.pdf')
    plt.show()
    plt.close()
