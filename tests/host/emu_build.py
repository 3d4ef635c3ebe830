"""TEST INFRASTRUCTURE: turns CUDA sources of the product into host C++ for the CPU execution harness (cuda_emu.h).

The product sources carry no emulation hooks: they are rewritten here, textually, into a generated .cpp:

  * `kernel<<<grid, block, smem, stream>>>(args);`  ->  `dq_emu::launch(grid, block, smem, [&]() { kernel(args); });`
  * `extern __shared__ T name[];`                   ->  `T* name = reinterpret_cast<T*>(DQ_EMU_DYNAMIC_SMEM);`
  * everything from a cut marker on is dropped (the tcgen05 / TMEM path cannot run on a CPU) and replaced by a tail; or
    (the bf16 training variant of tests/emu_qnet.py) only the regions the source marks as tcgen05 kernels are swapped for
    plain-loop stand-ins with the same launch arguments
  * includes of CUDA headers resolve to empty stubs in tests/host/emu_include/

Nothing here is used, timed or shipped by the product.
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "deepq_decoding_b200", "csrc")
SHIM = os.path.join(HERE, "cuda_emu.h")
STUBS = os.path.join(HERE, "emu_include")
GEN = os.path.join(HERE, "_gen")


def _match(text, i, open_ch, close_ch):
    """index just past the bracket that closes the one at text[i]"""
    depth = 0
    for j in range(i, len(text)):
        if text[j] == open_ch:
            depth += 1
        elif text[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced %s at %d" % (open_ch, i))


def _split_top(s):
    """split on commas that are not nested in (), <>, [] or {}"""
    out, depth, cur = [], 0, ""
    for k, ch in enumerate(s):
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>" and not (ch == ">" and k > 0 and s[k - 1] == "-"):        # `->` is not a bracket
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(text):
    out, pos = "", 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            return out + text[pos:]
        # kernel expression: identifier, optionally followed by a template argument list, right before <<<
        j = k
        if text[j - 1] == ">":
            depth, j = 0, k - 1
            while True:
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        m = re.search(r"[A-Za-z_][A-Za-z_0-9:]*$", text[:j])
        start = m.start()
        kern = text[start:k]
        e = text.find(">>>", k)
        cfg = _split_top(text[k + 3:e])
        while len(cfg) < 3:
            cfg.append("0")
        a0 = e + 3
        while text[a0] in " \n\t":
            a0 += 1
        assert text[a0] == "(", "launch of %s without argument list" % kern
        a1 = _match(text, a0, "(", ")")
        args = text[a0 + 1:a1 - 1]
        out += text[pos:start] + "dq_emu::launch(%s, %s, %s, [&]() { %s(%s); })" % (cfg[0], cfg[1], cfg[2], kern, args)
        pos = a1


def transform(path, cut_marker=None, tail="", swaps=()):
    """swaps: (begin marker, end marker, replacement) triples; the k-th marked region becomes the k-th replacement."""
    text = open(path).read()
    pos = 0
    for begin, end, repl in swaps:
        i = text.index(begin, pos)
        j = text.index(end, i) + len(end)
        text = text[:i] + repl + text[j:]
        pos = i + len(repl)
    if cut_marker:
        text = text[:text.index(cut_marker)] + "\n" + tail + "\n"
    text = re.sub(r'#include\s+"dq_ptx\.cuh"', "", text)
    text = text.replace('"../../include/dq_decoding.h"', '"%s"' % os.path.join(ROOT, "include", "dq_decoding.h"))
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?((?:unsigned\s+)?[A-Za-z_][\w:]*)\s+(\w+)\[\];",
                  r"\1* \2 = reinterpret_cast<\1*>(DQ_EMU_DYNAMIC_SMEM);", text)
    return rewrite_launches(text)


def build(out, sources, extra_flags=(), deps=()):
    """sources: list of (path, cut_marker, tail) tuples, or plain paths for sources that need no cut."""
    os.makedirs(GEN, exist_ok=True)
    all_deps = [SHIM, os.path.abspath(__file__), os.path.join(ROOT, "include", "dq_decoding.h"),
                os.path.join(CSRC, "dq_lattice.cuh"), os.path.join(CSRC, "dq_adam.cuh"), *deps]
    files = []
    for s in sources:
        path = s if isinstance(s, str) else s[0]
        all_deps.append(path)
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(f) for f in all_deps):
        return out
    for s in sources:
        path, cut, tail, *rest = (s, None, "") if isinstance(s, str) else s
        gen = os.path.join(GEN, os.path.basename(out).replace(".so", "_") + os.path.basename(path).replace(".cu", "_emu.cpp"))
        with open(gen, "w") as f:
            f.write('#line 1 "%s"\n' % path)
            f.write(transform(path, cut, tail, rest[0] if rest else ()))
        files.append(gen)
    cmd = ["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DDQ_EMU", "-include", SHIM,
           "-fsanitize=alignment", "-fno-sanitize-recover=alignment", *(["-fsanitize=address"] if os.environ.get("DQ_EMU_ASAN") else []),      # x86 forgives misaligned loads, the GPU does not
           
           "-I", STUBS, "-I", CSRC, *extra_flags]
    for f in files:
        cmd += ["-x", "c++", f]
    subprocess.check_call(cmd + ["-o", out])
    return out
