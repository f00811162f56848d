"""Extracts the trained tensors of every model set the reference ships into
weights/<set>/weights.npz (the GPU box has no /root/reference).  Run here:
    python tools/extract_weights.py [/root/reference]
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wakeword_detection_b200 import weights as W  # noqa: E402

SETS = {
    "CRNN": ("tf_lite_models/CRNN", "CRNN"),
    "Wavenet": ("tf_lite_models/Wavenet", "Wavenet"),
    "CRNN_arik_original": ("wwdetect/CRNN/models/Arik_CRNN_data_original", "CRNN"),
    "CRNN_arik_nosilence": ("wwdetect/CRNN/models/Arik_CRNN_data_nosilence", "CRNN"),
    "CRNN_arik_nosilence_enhanced": ("wwdetect/CRNN/models/Arik_CRNN_data_nosilence_enhanced", "CRNN"),
}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    for name, (sub, typ) in SETS.items():
        d = os.path.join(ref, sub)
        if not os.path.exists(os.path.join(d, "filter.tflite")):
            # the Arik_* directories hold no filter.tflite; the filter is shared
            tmp = tempfile.mkdtemp()
            for f in ("encode.tflite", "detect.tflite"):
                os.symlink(os.path.join(d, f), os.path.join(tmp, f))
            os.symlink(os.path.join(ref, "tf_lite_models/CRNN/filter.tflite"), os.path.join(tmp, "filter.tflite"))
            d = tmp
        w = W.load_model_dir(d, typ)
        out = os.path.join(ROOT, "weights", name)
        os.makedirs(out, exist_ok=True)
        W.save_npz(w, os.path.join(out, "weights.npz"))
        print(name, sum(v.size for v in w.values()), "values")


if __name__ == "__main__":
    main()
