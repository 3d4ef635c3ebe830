"""Host-side expansion alone (no GPU): dq_unpack_observations_host on random packed rows of the C3 shape, checked against the numpy
helper and timed per instruction set (children: DQ_HOST_NO_AVX512 / DQ_HOST_NO_AVX2) and per thread count (DQ_HOST_THREADS)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    from deepq_decoding_b200 import _lib
    from deepq_decoding_b200.envs import unpack_observations
    L = _lib.lib()
    n, d, ch = int(os.environ.get("DQ_N", "16384")), int(os.environ.get("DQ_D", "5")), int(os.environ.get("DQ_C", "7"))
    side = 2 * d + 1
    P, PW = side * side, (side * side + 63) // 64
    rng = np.random.default_rng(1)
    rows = rng.integers(0, 2 ** 63, size=(ch * PW, n), dtype=np.int64).view(np.uint64)
    rows[PW - 1::PW] &= np.uint64((1 << (P - 64 * (PW - 1))) - 1)
    obs = np.zeros((n, ch, side, side), dtype=np.uint8)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    _lib.check(L.dq_unpack_observations_host(vp(rows), n, n, d, ch, vp(obs)))
    ok = bool(np.array_equal(obs, unpack_observations(rows, n, d, ch)))
    for _ in range(20):
        L.dq_unpack_observations_host(vp(rows), n, n, d, ch, vp(obs))
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        L.dq_unpack_observations_host(vp(rows), n, n, d, ch, vp(obs))
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print("HOSTEXP " + json.dumps({"equal_to_numpy": ok, "median_us": ts[len(ts) // 2] * 1e6, "p90_us": ts[int(len(ts) * 0.9)] * 1e6,
                                    "GBps_median": obs.nbytes / ts[len(ts) // 2] / 1e9}))


if __name__ == "__main__":
    if os.environ.get("DQ_CHILD"):
        child()
    else:
        out = []
        for isa, extra in (("avx512", {}), ("avx2", {"DQ_HOST_NO_AVX512": "1"}), ("portable", {"DQ_HOST_NO_AVX2": "1"})):
            for th in ("default", "4", "1"):
                envv = dict(os.environ, DQ_CHILD="1", **extra)
                if th != "default":
                    envv["DQ_HOST_THREADS"] = th
                r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=envv, capture_output=True, text=True, timeout=300)
                line = [l for l in r.stdout.splitlines() if l.startswith("HOSTEXP ")]
                res = json.loads(line[-1][8:]) if line else {"error": (r.stderr or r.stdout)[-300:]}
                out.append(dict(res, isa=isa, threads=th))
        print(json.dumps({"cpus": len(os.sched_getaffinity(0)), "runs": out}, indent=1))
