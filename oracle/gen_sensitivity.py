"""gen_sensitivity.py -- how reproducible is the REFERENCE's own mask under rounding-level input noise?

For every golden node and horizon (30 / 100 epochs) the bit-exact port of the reference (oracle/gnnx_oracle.explain_dense_torch:
rel-L2 0.0 against the unmodified reference on the golden set, tests/test_oracle.py) is run again with EVERY entry of the initial
mask M0 moved by +-1 ulp (twelve random sign patterns; four more that also move every model weight by +-1 ulp) -- the size of
perturbation a different floating-point summation order introduces at every step.  spread = the largest relative L2 distance of those runs from the reference's golden mask.  A node with
spread <= 3e-5 is reproducible: any correct fp32 implementation must land within 1e-4 of the reference there.  A node with a
larger spread is chaotic (a ReLU kink crossed an epoch earlier or later): the reference cannot reproduce ITSELF to 1e-4 there.

tests/test_gpu_parity.py and tests/test_gpu_stream.py assert, per node:  rel-L2 <= max(1e-4, 3 x spread).
Writes tests/golden/<name>_sens.npz (nodes, spread_e30, spread_e100).  Needs only the committed fixtures; deterministic."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import gnnx_oracle as O  # noqa: E402
import util  # noqa: E402

torch.set_num_threads(1)


def one(args):
    name, node = args
    fx = util.load_fixture(name)
    gold = {30: np.load(util.GOLDEN + "/%s_golden_e30.npz" % name), 100: fx.gold}
    idx, srp, scol, sfeat, slabel, nbrs = O.extract_neighborhood(fx.rowptr, fx.col, fx.feat, fx.label, node, 3)
    n = len(nbrs)
    A = O.dense_from_csr(srp, scol)
    ei, ej = np.nonzero(A)
    m0 = fx.gold["n%d_m0" % node].astype(np.float32)
    pl = fx.pred_label[nbrs]
    out = []
    for ep in (30, 100):
        ref = gold[ep]["n%d_mask" % node]
        worst = 0.0
        for seed in range(16):
            rng = np.random.default_rng(1000 * seed + node)
            nudge = lambda x: np.where(rng.integers(0, 2, x.shape).astype(bool), np.nextafter(x, np.float32(np.inf)), np.nextafter(x, np.float32(-np.inf))).astype(np.float32)
            M0 = np.ones((n, n), np.float32)      # off-edge entries never reach the returned mask
            M0[ei, ej] = nudge(m0)
            W = fx.weights if seed < 12 else {k: nudge(np.asarray(v, np.float32)) for k, v in fx.weights.items()}
            res = O.explain_dense_torch(A, sfeat, slabel[idx], pl, idx, W, M0, hp=O.default_hparams(num_epochs=ep))
            worst = max(worst, O.rel_l2(np.asarray(res)[ei, ej], ref))
        out.append(worst)
    return node, out[0], out[1]


def main():
    import multiprocessing as mp
    for name in ["rand", "syn4", "syn1"]:
        fx = util.load_fixture(name)
        with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
            rows = pool.map(one, [(name, node) for node in fx.nodes])
        nodes = np.array([r[0] for r in rows]); s30 = np.array([r[1] for r in rows]); s100 = np.array([r[2] for r in rows])
        np.savez_compressed(os.path.join(util.GOLDEN, name + "_sens.npz"), nodes=nodes, spread_e30=s30, spread_e100=s100)
        print("%s: %d nodes; spread > 3e-5 at 30 epochs: %s ; at 100 epochs: %s" % (
            name, len(nodes), {int(n): "%.1e" % s for n, s in zip(nodes, s30) if s > 3e-5}, {int(n): "%.1e" % s for n, s in zip(nodes, s100) if s > 3e-5}), flush=True)


if __name__ == "__main__":
    main()
