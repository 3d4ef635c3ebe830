#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` capture of the env-step kernel: warp-instructions executed and
stall samples summed by the line of csrc/dq_env.cu / dq_lattice.cuh the SASS came from (line table of the SAME build), grouped by the
warp role that owns the line.    python tools/ncu_lines.py <capture.ncu-rep> [lib.so] [--top 40]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_source_breakdown as T  # noqa: E402


def role_ranges(src):
    """line ranges of the three roles and the common parts, from the markers in the kernel source"""
    lines = open(src).read().splitlines()
    find = lambda s, a=0: next(i + 1 for i, l in enumerate(lines) if i >= a and s in l)
    k0 = find("env_step_kernel(const EnvParams p")
    phys = find("PHYSICS: lane = lattice")
    wr = find("WRITERS", phys)
    gen = find("GENERATORS", wr)
    tail = find("the last step's observation bytes, by every thread")
    end = find("Uniform pick over the sorted legal actions")
    # helper functions defined before the kernel, by the role that calls them
    rec = find("void record_event(")
    draw = find("void draw_flip_masks(")
    roll = find("struct Rollout {")
    place = find("void place_layer(")
    legal = find("void legal_words(")
    wobs = find("void write_observations_unaligned(")
    ref = find("int referee_class(")
    sync = find("void bar_sync_named(")
    return [(ref - 3, sync - 2, "physics (referee lookup)"), (rec - 6, roll, "generators (draw_flip_masks / record_event / generate_attempt)"),
            (place - 5, legal - 4, "writers (place_layer / expand16 / stores)"), (legal - 4, wobs - 5, "physics (legal_words)"),
            (wobs - 5, k0 - 2, "writers (write_observations)"),
            (k0 - 2, phys, "prologue"), (phys, wr, "physics"), (wr, gen, "writers"), (gen, tail, "generators"), (tail, end, "epilogue")], k0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("lib", nargs="?", default="deepq_decoding_b200/libdq_decoding.so")
    ap.add_argument("--source", default="deepq_decoding_b200/csrc/dq_env.cu", help="the dq_env.cu the captured library was built from")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--kernel", default="env_step_kernelILi5ELb0")
    a = ap.parse_args()
    seq = T.line_table(a.lib, a.kernel)
    hdr, rows = T.source_page(a.rep)
    assert len(seq) == len(rows), "capture and library are different builds (%d vs %d instructions)" % (len(rows), len(seq))
    ci = {n: i for i, n in enumerate(hdr)}
    ranges, k0 = role_ranges(a.source)
    per_line, per_role = collections.defaultdict(lambda: [0, 0, 0]), collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for (addr, loc, op), r in zip(seq, rows):
        inst, samp, nis = int(r[ci["Instructions Executed"]] or 0), int(r[ci["Warp Stall Sampling (All Samples)"]] or 0), int(r[ci["Warp Stall Sampling (Not-issued Samples)"]] or 0)
        f, l = loc if loc else ("?", 0)
        role = f
        if f == "dq_env.cu":
            role = next((n for lo, hi, n in ranges if lo <= l < hi), "helpers (%s)" % ("before kernel" if l < k0 else "after"))
        for acc in (per_line[(f, l)], per_role[role], tot):
            acc[0] += inst; acc[1] += samp; acc[2] += nis
    print("| owner (by source line) | warp-instructions | % | stall samples | % |\n|---|---|---|---|---|")
    for k, v in sorted(per_role.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %d | %.1f |" % (k, v[0], 100.0 * v[0] / max(1, tot[0]), v[1], 100.0 * v[1] / max(1, tot[1])))
    print("\n| file:line | warp-instructions | stall samples | % of samples |\n|---|---|---|---|")
    for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print("| %s:%d | %d | %d | %.1f |" % (k[0], k[1], v[0], v[1], 100.0 * v[1] / max(1, tot[1])))


if __name__ == "__main__":
    main()
