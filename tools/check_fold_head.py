"""GPU check of the opt-in folded head of the bf16 acting path (DQ_QNET_FOLD_HEAD=1: Dense(A) + dueling head as one affine map
staged by dq_qnet_prepare_tc): Q values against the unfolded bf16 path and the fp32 path on the published d5_dp/0.007 agent,
and the forward time of both.  The flag is read once per process, so each arm runs in its own process.

    python tools/check_fold_head.py            # prints one JSON line
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def arm(out):
    import numpy as np
    import torch
    from deepq_decoding_b200 import agents as A
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    import ctypes as C
    from deepq_decoding_b200 import _lib
    n = 16384
    env = VecSurfaceCodeEnv(5, 0.007, 0.007, "DP", False, 5, None, n_envs=n, seed=5)
    env.reset()
    z = np.load(os.path.join(ROOT, "tests", "golden", "dqn_d5_dp_0.007.npz"))
    ag = A.DQNAgent(model=A.build_convolutional_nn([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], (7, 11, 11), env.num_actions),
                    nb_actions=env.num_actions, memory=A.SequentialMemory(limit=100), test_policy=A.GreedyQPolicy(masked_greedy=True),
                    enable_dueling_network=True, device=env.device, act_precision="bf16")
    ag.compile(A.Adam(lr=1e-5), max_envs=n)
    ag.model.set_keras_weights([(z["conv%d_k" % i], z["conv%d_b" % i]) for i in range(3)], [(z["dense%d_k" % i], z["dense%d_b" % i]) for i in range(3)])
    for t in range(10):                                   # a few steps so that the boards are not all fresh volumes
        env.step(env.random_legal_actions(t))
    rows, nrows, stride = C.c_void_p(), C.c_int64(), C.c_int64()
    _lib.check(_lib.lib().dq_env_packed_obs(env._h, C.byref(rows), C.byref(nrows), C.byref(stride)))
    q16 = ag.model.forward_packed(rows.value, stride.value, n, precision="bf16").clone()
    q32 = ag.model.forward_packed(rows.value, stride.value, n).clone()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        ag.model.forward_packed(rows.value, stride.value, n, precision="bf16")
    b.record()
    torch.cuda.synchronize()
    np.savez(out, q16=q16.cpu().numpy(), q32=q32.cpu().numpy(), ms=a.elapsed_time(b) / 50)


def main():
    if len(sys.argv) > 1:
        return arm(sys.argv[1])
    import numpy as np
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for name, flag in (("unfolded", "0"), ("folded", "1")):
            out = os.path.join(d, name + ".npz")
            subprocess.check_call([sys.executable, os.path.abspath(__file__), out], env=dict(os.environ, DQ_QNET_FOLD_HEAD=flag))
            z = np.load(out)
            res[name] = {"forward_ms": float(z["ms"]), "max_abs_diff_vs_fp32": float(np.abs(z["q16"] - z["q32"]).max()),
                         "greedy_agreement_with_fp32": float((z["q16"].argmax(1) == z["q32"].argmax(1)).mean())}
            res[name + "_q"] = z["q16"]
        res["max_abs_diff_folded_vs_unfolded"] = float(np.abs(res.pop("folded_q") - res.pop("unfolded_q")).max())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
