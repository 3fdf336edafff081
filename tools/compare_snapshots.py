#!/usr/bin/env python
"""Compare two PSTJ trajectory snapshots of the same scene (include/ps_snapshot.h), frame by frame.

Meant for the step DESIGN.md §2 leaves open: a file written by the headless .NET driver of INTEGRATION.md §6 (the real
reference) against one written here by the oracle or the CUDA path.  Prints per frame whether the states are bit
identical, the largest relative difference per field over dynamic bodies, and whether the ordered candidate lists,
the colliding-pair sets and the contact-feature ids agree; then the verdict against the stated tolerance policy
(SURVEY.md §8(c)): sets and ids identical until the first discrete divergence, states within --rtol-step per step
where the previous frame was still within tolerance.

    python tools/compare_snapshots.py reference.pstj ours.pstj [--rtol 1e-9]
    python tools/compare_snapshots.py --record-cuda tests/golden/snap_c2_two_worlds.pstj out.pstj   # replay on the GPU
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from physicssimulator_b200 import snapshot  # noqa: E402


def record_cuda(src: str, dst: str) -> None:
    from physicssimulator_b200 import _native
    from physicssimulator_b200.configuration import ps_step_config
    gold = snapshot.load(src)
    cfg = ps_step_config.from_buffer_copy(bytes(gold.config))
    cfg.record_contacts = 1
    batch = _native.Batch(gold.bodies(0), gold.world_offset, cfg, device=0)
    try:
        t = snapshot.record(gold.bodies(0), gold.world_offset, gold.config, [f.step for f in gold.frames[1:]], batch.step,
                            batch.get_state, batch.contacts, batch.candidates, producer=snapshot.PRODUCER_CUDA, scene=gold.scene)
    finally:
        batch.close()
    snapshot.save(dst, t)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("a")
    ap.add_argument("b")
    ap.add_argument("--rtol", type=float, default=1e-9, help="relative tolerance on states while the discrete decisions agree")
    ap.add_argument("--record-cuda", action="store_true", help="replay snapshot A on the GPU through the C ABI and write it to B")
    args = ap.parse_args()
    if args.record_cuda:
        record_cuda(args.a, args.b)
        args.a, args.b = args.a, args.b
    A, B = snapshot.load(args.a), snapshot.load(args.b)
    print(f"A: {args.a}: {snapshot.PRODUCER_NAMES.get(A.producer, '?')}, scene '{A.scene}', {A.n_worlds} worlds, {A.n_bodies} bodies, {len(A.frames)} frames")
    print(f"B: {args.b}: {snapshot.PRODUCER_NAMES.get(B.producer, '?')}, scene '{B.scene}', {B.n_worlds} worlds, {B.n_bodies} bodies, {len(B.frames)} frames")
    diffs = snapshot.compare(A, B)
    diverged_at = None
    worst = 0.0
    ok = True
    print(f"{'step':>6} {'bits':>5} {'pos':>9} {'vel':>9} {'rot':>9} {'ang_mom':>9} {'cand':>5} {'pairs':>5} {'feat':>5} {'contact abs':>11}")
    for d in diffs:
        f = lambda v: "-" if v is None else ("yes" if v else "NO")  # noqa: E731
        print(f"{d.step:6d} {f(d.bit_exact):>5} {d.max_rel['pos']:9.2e} {d.max_rel['vel']:9.2e} {d.max_rel['rot']:9.2e} {d.max_rel['ang_mom']:9.2e} "
              f"{f(d.candidates_equal):>5} {f(d.pairs_equal):>5} {f(d.features_equal):>5} {'-' if d.contact_max_abs is None else format(d.contact_max_abs, '11.2e')}")
        discrete_ok = all(v is None or v for v in (d.candidates_equal, d.pairs_equal, d.features_equal))
        if diverged_at is None and not discrete_ok:
            diverged_at = d.step
        if diverged_at is None:
            m = max(d.max_rel.values())
            worst = max(worst, m)
            ok = ok and m <= args.rtol
    if diverged_at is None:
        print(f"verdict: discrete decisions identical in every frame; largest relative state difference {worst:.3e} "
              f"({'within' if ok else 'ABOVE'} rtol {args.rtol:g})")
    else:
        print(f"verdict: first discrete divergence (candidates / colliding pairs / feature ids) at step {diverged_at}; "
              f"before it the largest relative state difference was {worst:.3e} ({'within' if ok else 'ABOVE'} rtol {args.rtol:g})")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
