#!/usr/bin/env python
"""Per-phase breakdown of the env-step kernel from an `ncu --set full --import-source on` capture.

    python tools/ncu_source_breakdown.py gpurun_out/v8_rollout64.ncu-rep deepq_decoding_b200/libdq_decoding.so \
        --tiles 1024 --steps 64 > profiles/<name>.md

Joins the capture's source page (per SASS instruction: warp-instructions executed, stall samples and their reasons) with the
line table of the SAME build (nvdisasm -g on the library's cubin: the capture and the library must be one build, checked opcode by
opcode) and sums both by phase of the step: the two block barriers of the kernel split the SASS into the window of phase A / D, phase B
and phase C; within a window the source line says whose code it is.  Numbers are per tile-step (one CTA, one step of its lattices).
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

KERNEL = "env_step_kernelILi5ELb0"


def line_table(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith("dq_env.") and f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    seq, cur, on = [], None, False
    for ln in text.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            seq.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return seq


def source_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    return rows[hi], rows[hi + 1:]


def opcode(t):
    parts = t.split()
    return (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]


def owner(f, l):
    if f == "dq_env.cu":
        for lo, hi, name in OWNERS:
            if lo <= l <= hi:
                return name
    if f == "dq_lattice.cuh":
        for lo, hi, name in LATTICE:
            if lo <= l <= hi:
                return name
    return "%s:%d" % (f, l)


# Owners of source lines.  FIXED_* are the ranges at commit af650fa (the v8 kernel the round-1 captures were taken from); for a library
# built from a later source pass --source <dir of that source> and the ranges are derived from the function signatures and the
# phase comments of dq_env.cu / dq_lattice.cuh themselves (derive_owners).
FIXED_OWNERS = [(98, 106, "referee table lookup"), (108, 122, "record_event (fired draws)"), (124, 171, "draw_flip_masks (screen + walk)"),
                (173, 215, "generate_volume after the draws"), (244, 270, "stream gather"), (272, 276, "expand16"),
                (278, 293, "legal_words"), (295, 333, "write_observations"), (335, 403, "loop control / prologue"),
                (405, 426, "helpers: policy words, D call"), (428, 506, "phase A body (warp 0)"), (507, 519, "barrier 1 / light cell"),
                (520, 581, "phase B task body"), (782, 802, "phase C / tail")]
FIXED_LATTICE = [(34, 41, "popc64"), (42, 112, "Lat<D> geometry"), (113, 120, "true_syndrome"), (121, 125, "homology_label"),
                 (126, 139, "qubit grid <-> compact"), (140, 160, "stabilizer grid <-> compact"), (161, 171, "joint referee index"),
                 (172, 180, "split referee index"), (181, 192, "adjacent / neighbour qubits"), (193, 215, "syndrome_layer_bitmap"),
                 (216, 227, "action_layer_bitmap"), (228, 250, "philox4x32_10"), (251, 267, "extract / popc32"), (268, 290, "select64")]
OWNERS, LATTICE = FIXED_OWNERS, FIXED_LATTICE

ENV_MARKS = [("int lut2(", "referee table lookup"), ("void record_event(", "record_event (fired draws)"),
             ("bool draw_flip_masks(", "draw_flip_masks (screen + walk)"), ("u64 generate_volume(", "generate_volume after the draws"),
             ("struct Rollout", None), ("u32 gather32(", "stream gather"), ("uint4 expand16(", "expand16"), ("void store_obs16(", "write_observations"),
             ("void named_barrier(", "render_dirty (deferred bitmaps)"), ("u32 fresh_word(", "fresh_word (deferred stream spans)"),
             ("void legal_words(", "legal_words"), ("void write_observations_unaligned(", "write_observations"),
             ("env_step_kernel(const EnvParams p", "loop control / prologue"), ("// ---- phase A", "helpers: policy words, D call"),
             ("        const int e = env0 + lane;", "phase A body (warp 0)"), ("a light step's action lights", "barrier 1 / light cell"),
             ("// ---- phase B", "phase B task body"), ("} else if constexpr (DQ_BATCHB == 2)", "phase B (batched finalisation)"),
             ("// DQ_BATCHB=1.  Everything after the draws", "phase B (batched)"), ("// ---- phase C", "phase C / tail"),
             ("__global__ void policy_random_legal_kernel(", None)]
LAT_MARKS = [("int popc64(", "popc64"), ("struct Lat", "Lat<D> geometry"), ("u64 plaquette_parity(", "true_syndrome"), ("int homology_label(", "homology_label"),
             ("u64 qubits_compact_to_grid(", "qubit grid <-> compact"), ("u64 stabs_grid_to_compact(", "stabilizer grid <-> compact"),
             ("u32 stabs_grid_to_joint_index(", "joint referee index"), ("u32 stabs_grid_to_type_index(", "split referee index"),
             ("u64 qubits_adjacent_to(", "adjacent / neighbour qubits"), ("u32 spread2_8(", "syndrome_layer_bitmap"),
             ("void action_layer_bitmap(", "action_layer_bitmap"), ("u32 mulhi32(", "philox4x32_10"), ("u64 extract_bits(", "extract / popc32"),
             ("int select64(", "select64")]


def derive_owners(path, marks):
    """[(first line, last line, owner)] from the first line that contains each marker, in file order; a None owner ends a range."""
    lines = open(path).read().splitlines()
    found = []
    for text, name in marks:
        for i, ln in enumerate(lines, 1):
            if text in ln and (not found or i > found[-1][0]):
                found.append((i, name))
                break
    found.sort()
    out = []
    for k, (i, name) in enumerate(found):
        end = found[k + 1][0] - 1 if k + 1 < len(found) else len(lines)
        if name:
            out.append((i, end, name))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("lib")
    ap.add_argument("--tiles", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--kernel", default=KERNEL)
    ap.add_argument("--source", default=None, help="directory of the dq_env.cu / dq_lattice.cuh the library was built from (default: the af650fa line ranges)")
    a = ap.parse_args()
    if a.source:
        global OWNERS, LATTICE
        OWNERS = derive_owners(os.path.join(a.source, "dq_env.cu"), ENV_MARKS)
        LATTICE = derive_owners(os.path.join(a.source, "dq_lattice.cuh"), LAT_MARKS)
    seq = line_table(a.lib, a.kernel)
    hdr, data = source_page(a.rep)
    assert len(seq) == len(data), "capture and library are different builds (%d vs %d instructions)" % (len(data), len(seq))
    assert all(opcode(t) == opcode(r[1].strip()) for (_, _, t), r in zip(seq, data)), "capture and library are different builds"
    ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    bars = [ad for ad, _, t in seq if "BAR.SYNC" in t]
    assert len(bars) == 4, "expected 2 prologue barriers + 2 per step"
    last_exit = max(ad for ad, _, t in seq if t.endswith("EXIT"))
    names = ["prologue", "prologue", "window A / D", "phase B", "phase C / tail", "out-of-line functions"]

    def window(ad):
        for k, b in enumerate(bars):
            if ad <= b:
                return names[k]
        return names[4] if ad <= last_exit else names[5]

    per = a.tiles * a.steps
    tot_i = sum(int(r[ci]) for r in data)
    tot_s = sum(int(r[cs]) for r in data)
    agg, wagg, reasons = collections.OrderedDict(), collections.OrderedDict(), collections.Counter()
    for (ad, c, t), r in zip(seq, data):
        w = window(ad)
        k = (w, owner(*c))
        wait = "BSSY" in t and int(r[stalls[0][0]]) > 0.9 * max(int(r[cs]), 1) and int(r[cs]) > 100     # samples parked at a block barrier
        if wait:
            k = (w, "WAITING at the block barrier before this window")
        e = agg.setdefault(k, [0, 0, 0, 0])
        e[0] += int(r[ci]); e[1] += int(r[cs]); e[2] += 1; e[3] += int(r[ct])
        e = wagg.setdefault(w, [0, 0])
        e[0] += int(r[ci]); e[1] += int(r[cs])
        for i, h in stalls:
            reasons[h] += int(r[i])
    print("# env_step_kernel<5,false>: where the instructions and the warp time go\n")
    print("Source: `%s` joined with the line table of the same build; %d tiles x %d steps; %d warp-instructions (%.0f per tile-step, "
          "%.0f per lattice-step at 16 lattices per tile); %d stall samples.\n" % (os.path.basename(a.rep), a.tiles, a.steps, tot_i,
                                                                                   tot_i / per, tot_i / per / 16, tot_s))
    print("| window | warp-instructions per tile-step | share of instructions | share of warp time (samples) |\n|---|---|---|---|")
    for w, (i, s) in wagg.items():
        print("| %s | %.0f | %.1f %% | %.1f %% |" % (w, i / per, 100 * i / tot_i, 100 * s / tot_s))
    print("\n| window | code | static SASS | warp-instructions per tile-step | active lanes | share of warp time |\n|---|---|---|---|---|---|")
    for (w, k), (i, s, n, th) in agg.items():
        if i / per >= 4 or s / tot_s >= 0.004:
            print("| %s | %s | %d | %.1f | %.1f | %.1f %% |" % (w, k, n, i / per, th / max(i, 1), 100 * s / tot_s))
    print("\nStall reasons over all samples: " + ", ".join("%s %.1f %%" % (h, 100 * n / tot_s) for h, n in reasons.most_common(9)) + ".\n")
    print("Most-sampled instructions outside the barrier waits:\n\n| share of warp time | line | instruction | top reasons |\n|---|---|---|---|")
    rows = sorted(((int(r[cs]), c, t, r) for (ad, c, t), r in zip(seq, data) if "BSSY" not in t), key=lambda x: -x[0])[:12]
    for s, c, t, r in rows:
        top = sorted(((int(r[i]), h) for i, h in stalls if int(r[i]) > 0), reverse=True)[:2]
        print("| %.2f %% | %s:%d | `%s` | %s |" % (100 * s / tot_s, c[0], c[1], t[:60], ", ".join("%s %d" % (h, n) for n, h in top)))


if __name__ == "__main__":
    main()
