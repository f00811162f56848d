"""Drop-in for the reference's clip-level evaluation of the TF-Lite models
(reference: utils/evaluate_tf_lite_opts.py:16-131).

Same functions, arguments and return values: `load_tf_models`, `load_data`, `models_predict`,
`metrics`, `main` with the reference's flags.  `models_predict` runs all N clips as ONE batch
through the CUDA library instead of N batch-1 interpreter invocations; the decision is the
reference's non-strict `posterior >= threshold` (:65), unlike the trigger's strict `>`.

Differences forced by this image: `h5py` is not installable here, so `load_data` also accepts the
npz twin written by `save_npz` (one `[T, 40]` array per key + `<key>/is_hotword`); with h5py
present the reference's H5 layout (`h5[key][()]`, `attrs['is_hotword']`, :35-47) is read as is.
"""
from __future__ import annotations

import argparse
import os
import pickle
import sys
import time
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .models import TFLiteModel

_LABEL = "/is_hotword"


def load_tf_models(models_dir, quant=False, derive_quant=False):
    """evaluate_tf_lite_opts.py:16-33.  quant=True opens `encode-quant.tflite` / `detect-quant.tflite` (weights stored as
    float16, float32 arithmetic); with derive_quant=True a directory that only holds the float32 files gets the variant
    derived by the converter's rounding instead of an error."""
    suffix = "-quant" if quant else ""
    encode_model = TFLiteModel(model_path=os.path.join(models_dir, "encode%s.tflite" % suffix), derive_quant=derive_quant)
    detect_model = TFLiteModel(model_path=os.path.join(models_dir, "detect%s.tflite" % suffix), derive_quant=derive_quant)
    return encode_model, detect_model


def save_npz(data_file: str, clips: Dict[str, np.ndarray], labels: Dict[str, int]) -> None:
    """npz twin of the H5 feature file (utils/filter_dataset_to_h5.py writes `[T, 40]` + attrs)."""
    out = {}
    for k, v in clips.items():
        out[k] = np.asarray(v, np.float32)
        out[k + _LABEL] = np.uint8(labels[k])
    np.savez(data_file, **out)


def load_data(data_file, timesteps, num_features) -> Tuple[np.ndarray, np.ndarray]:
    """evaluate_tf_lite_opts.py:35-47: every clip truncated / zero-padded to `timesteps` rows;
    keys in the file's own order; labels uint8."""
    labels: List[int] = []
    if str(data_file).endswith(".npz"):
        with np.load(data_file) as z:
            keys = [k for k in z.files if '/' not in k]      # '<key>/<attr>' entries are the attributes
            X = np.zeros((len(keys), timesteps, num_features), dtype=np.float32)
            for i, key in enumerate(keys):
                labels.append(int(z[key + _LABEL]))
                features = z[key][:timesteps]
                X[i, :features.shape[0], :features.shape[1]] = features
    else:
        try:
            import h5py
        except ImportError as e:
            raise ImportError("reading %s needs h5py (not in this image); use the npz twin (save_npz)" % data_file) from e
        with h5py.File(data_file, 'r') as h5:
            keys = list(h5.keys())
            X = np.zeros((len(keys), timesteps, num_features), dtype=np.float32)
            for i, key in enumerate(keys):
                labels.append(h5[key].attrs['is_hotword'])
                features = h5[key][()][:timesteps]
                X[i, :features.shape[0], :features.shape[1]] = features
    return X, np.array(labels, dtype=np.uint8)


def posteriors(encode_model, detect_model, X, model_type, batch: int = 16384) -> np.ndarray:
    """Wake-class probability of every clip, `[N]` float32 (the values `models_predict` thresholds)."""
    eng = encode_model._engine
    if detect_model._engine is not eng:
        raise ValueError("encode and detect models come from different model directories")
    kind = "CRNN" if encode_model.input_details[0]["shape"].shape[0] == 4 else "Wavenet"
    if model_type != kind:
        # the reference would feed a wrongly shaped tensor to the interpreter (:56-62)
        raise ValueError("Cannot set tensor: Dimension mismatch. model_type %s but the models are %s" % (model_type, kind))
    X = np.asarray(X)
    if X.dtype != np.float32:
        raise ValueError("Cannot set tensor: Got value of type %s but expected type FLOAT32" % X.dtype)
    if X.ndim != 3 or X.shape[1] != eng.L or X.shape[2] != eng.n_mel:
        raise ValueError("Cannot set tensor: Dimension mismatch. Got %s but expected [N, %d, %d]"
                         % (list(X.shape), eng.L, eng.n_mel))
    torch = eng.torch
    out = np.empty((X.shape[0],), np.float32)
    for b0 in range(0, X.shape[0], batch):
        x = torch.from_numpy(np.ascontiguousarray(X[b0:b0 + batch])).to(eng.device)
        out[b0:b0 + batch] = eng.posteriors(x, hop=1)[:, 0].cpu().numpy()   # n_frames == L: one window per clip
    return out


def models_predict(encode_model, detect_model, X, model_type, threshold=0.5) -> List[int]:
    """evaluate_tf_lite_opts.py:49-67.  The reference indexes the detect output at `[0][0][1]`; the
    wake-class probability is the last output for either detect head."""
    post = posteriors(encode_model, detect_model, X, model_type)
    return [1 if posterior >= threshold else 0 for posterior in post]


def metrics(preds: Sequence[int], targets: Sequence[int]):
    """evaluate_tf_lite_opts.py:69-87 — including its naming: 'recall' is tp/(tp+fp) and 'precision'
    tp/(tp+fn) there; kept so that result files compare equal."""
    from sklearn.metrics import balanced_accuracy_score, confusion_matrix
    cf = confusion_matrix(targets, preds)
    print(f'\nConfusion matrix:\n {cf}')
    tn, fp, fn, tp = cf.ravel()
    results = {'true_negative': tn, 'false_positive': fp,
               'true_positive': tp, 'false_negative': fn,
               'recall': tp / (tp + fp), 'precision': tp / (tp + fn),
               'accuracy': balanced_accuracy_score(targets, preds)}
    for key, val in results.items():
        print(f'{key}: {val}')
    return results


def parse_args(argv=None):
    """evaluate_tf_lite_opts.py:89-99 (same flags and defaults)."""
    parser = argparse.ArgumentParser(description='Evaluation script for TF-Lite models.')
    parser.add_argument('--tf_models_dir', type=str, default='CRNN_tf_model', help='Directory to saved TF-Lite models')
    parser.add_argument('--dataset_dir', type=str, default='data_speech_isolated/silero', help='Directory with testing vectors in H5 format')
    parser.add_argument('--testset', type=str, default='test.h5', help='Filename for testing vectors in H5 (or npz) format')
    parser.add_argument('--timesteps', type=int, default=151, help='Number of timesteps used as input to models')
    parser.add_argument('--num_features', type=int, default=40, help='Number of features per timestep used as input to models')
    parser.add_argument('--model_type', type=str, default='CRNN', choices=['CRNN', 'Wavenet'],
                        help='Model type being evaluated.')
    return parser.parse_args(argv)


def main(args) -> int:
    """evaluate_tf_lite_opts.py:102-131: float32 and float16-weight variants of the model directory."""
    start = time.time()
    encode_model, detect_model = load_tf_models(args.tf_models_dir)
    X, y = load_data(os.path.join(args.dataset_dir, args.testset), args.timesteps, args.num_features)
    results = {}
    print(f'Testing {args.model_type} TF-Lite models with 32-bit floats')
    preds = models_predict(encode_model, detect_model, X, args.model_type)
    results['float32'] = metrics(preds, y)
    if not os.path.isfile(os.path.join(args.tf_models_dir, "encode-quant.tflite")):
        print('no encode-quant.tflite in %s: the float16 arm runs on the float32 weights rounded to float16 '
              '(what convert_*_tflite.py stores)' % args.tf_models_dir)
    encode_q, detect_q = load_tf_models(args.tf_models_dir, quant=True, derive_quant=True)
    print(f'Testing {args.model_type} TF-Lite models with 16-bit floats')
    preds = models_predict(encode_q, detect_q, X, args.model_type)
    results['float16'] = metrics(preds, y)
    pickle.dump(results, open(os.path.join(args.tf_models_dir, 'tf_lite_results.npy'), 'wb'))
    print(f'Script completed in {time.time()-start:.2f} secs')
    return 0


if __name__ == '__main__':
    sys.exit(main(parse_args()))
