"""K4 must give every board the same bits whatever its place in the batch and whatever else is in the batch (the reference-side shims compare
evaluations made in different batch compositions). Runs the forward on permuted / truncated batches and compares per board, bit for bit."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob

def main(blocks, filters, q, n, S=15):
    eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, S, S), max_boards=n, blocks=blocks, filters=filters, q_head=bool(q))
    eng.load_weights(netblob.pack(netblob.random_tensors(S, S, blocks, filters, bool(q), seed=5), S, S, blocks, filters, bool(q)))
    rng = np.random.default_rng(1)
    feats = rng.integers(0, 2**31 - 1, (n, S * S), dtype=np.int64).astype(np.uint32)
    base = eng.forward(feats, want_q=bool(q))
    bad = 0
    for trial in range(12):
        m = int(rng.integers(1, n + 1))
        perm = rng.permutation(n)[:m]
        out = eng.forward(np.ascontiguousarray(feats[perm]), want_q=bool(q))
        for a, b in zip(base, out):
            if a is None:
                continue
            diff = (a[perm].view(np.uint32) != b.view(np.uint32))
            if diff.any():
                bad += 1
                rows = np.unique(np.nonzero(diff.reshape(m, -1))[0])
                print(f"trial {trial} m={m}: {diff.sum()} differing words in {rows.size} boards, first rows {rows[:8]}, max abs diff {np.abs(a[perm]-b).max():.3g}")
    print(f"{blocks}x{filters} q={q} n={n}: {'OK' if bad == 0 else 'MISMATCH'}")
    eng.close()

if __name__ == "__main__":
    main(4, 64, 0, 300)
    main(4, 64, 0, 37)
    main(20, 128, 1, 700)
