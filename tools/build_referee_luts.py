"""Tabulates the reference's shipped referee MLPs into the packed tables the package ships.

Run in the build container (needs /root/reference):  python tools/build_referee_luts.py
Writes deepq_decoding_b200/data/referee_d5_{X,DP}.lut (+ min-weight tables for d=3, d=7).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepq_decoding_b200 import referee as R  # noqa

REF = "/root/reference/example_notebooks/referee_decoders"

if __name__ == "__main__":
    os.makedirs(R.DATA_DIR, exist_ok=True)
    for d in (3, 7):
        for model in ("X", "DP"):
            t = time.time()
            R.min_weight(d, model).save(os.path.join(R.DATA_DIR, "referee_d%d_%s.lut" % (d, model)))
            print("min-weight d=%d %s: %.1fs" % (d, model, time.time() - t), flush=True)
    for model in ("X", "DP"):
        t = time.time()
        lut = R.from_keras_mlp(os.path.join(REF, "nn_d5_%s_p5" % model), 5, model,
                               progress=lambda hi, n: (hi % (1 << 21) == 0) and print("  %d/%d" % (hi, n), flush=True))
        out = os.path.join(R.DATA_DIR, "referee_d5_%s.lut" % model)
        lut.save(out)
        print("d=5 %s: %.1fs -> %s (%d bytes)" % (model, time.time() - t, out, os.path.getsize(out)), flush=True)
