"""TEST INFRASTRUCTURE ONLY: builds tests/cuda_emul/libljgpu_emul.so - the product's host + device sources
(lj_sim.cu, lj_sort.cu, lj_host.cpp) compiled by g++ against the CUDA-on-CPU shim.  Kernel launches
`k<<<grid, block, smem, stream>>>(args)` are rewritten to `ljemul::launch_grid(grid, block, smem, [&]{ k(args); })`
in a scratch copy; the product sources are not modified.  The result exports the C ABI of include/ljgpu.h and can be
loaded by the Python binding through LJGPU_LIBRARY to check host logic and kernel logic together without a GPU.
It is never shipped and never loaded unless a test asks for it by path."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "lennard-jones-live-simulator_b200")
OUT = os.path.join(HERE, "libljgpu_emul.so")
SCRATCH = os.path.join(HERE, "_build")


def rewrite_launches(text):
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            break
        # kernel expression: identifier, optionally followed by <template args>, ending right before <<<
        s = k
        if text[s - 1] == ">":
            depth, s = 0, s - 1
            while True:
                if text[s] == ">":
                    depth += 1
                elif text[s] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                s -= 1
        while s > 0 and (text[s - 1].isalnum() or text[s - 1] in "_:"):
            s -= 1
        kernel = text[s:k]
        e = text.index(">>>", k)
        cfg = text[k + 3:e]
        # split the launch configuration on top-level commas
        parts, depth, cur = [], 0, ""
        for ch in cfg:
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur.strip()); cur = ""
            else:
                cur += ch
        parts.append(cur.strip())
        assert len(parts) == 4, (kernel, cfg)
        a = text.index("(", e)
        depth, b = 0, a
        while True:
            if text[b] == "(":
                depth += 1
            elif text[b] == ")":
                depth -= 1
                if depth == 0:
                    break
            b += 1
        args = text[a + 1:b]
        out.append(text[pos:s])
        out.append(f"ljemul::launch_grid({parts[0]}, {parts[1]}, {parts[2]}, [&] {{ {kernel}({args}); }})")
        pos = b + 1
    return "".join(out)


def build(force=False, asan=False):
    """asan=True: the same library with -fsanitize=address (memory checking of kernel and host logic where
    compute-sanitizer is not available); load it with LD_PRELOAD="libasan.so libstdc++.so.6" - see run_asan.sh."""
    OUT = os.path.join(HERE, "_build", "libljgpu_emul_asan.so") if asan else globals()["OUT"]
    SCRATCH = os.path.join(HERE, "_build", "asan") if asan else globals()["SCRATCH"]
    extra = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []
    srcs = [os.path.join(PKG, "csrc", f) for f in ("lj_sim.cu", "lj_sort.cu")]
    deps = srcs + [os.path.join(PKG, "csrc", f) for f in ("lj_common.cuh", "lj_kernels.cuh", "lj_tile.cuh", "lj_nccl.h")]
    deps += [os.path.join(PKG, "host", "lj_host.cpp"), os.path.join(ROOT, "include", "ljgpu.h"),
             os.path.join(HERE, "cuda_emul.h"), os.path.join(HERE, "cuda_runtime_emul.h"), os.path.join(HERE, "nccl_emul.cpp"), os.path.join(HERE, "emul_api.cpp"),
             os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.join(SCRATCH, "csrc"), exist_ok=True)
    objs = []
    for src in srcs:
        text = rewrite_launches(open(src).read())
        # the scratch copy sits two levels below a fake package root so that the relative includes still resolve
        dst = os.path.join(SCRATCH, "csrc", os.path.basename(src).replace(".cu", "_emul.cpp"))
        text = text.replace('#include "../../include/ljgpu.h"', f'#include "{os.path.join(ROOT, "include", "ljgpu.h")}"')
        open(dst, "w").write(text)
        obj = dst.replace(".cpp", ".o")
        subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-pthread", "-mfma", "-ffp-contract=off", "-DLJ_EMUL"] + extra +
                       ["-I", os.path.join(HERE, "include"), "-I", os.path.join(PKG, "csrc"),
                        "-include", os.path.join(HERE, "include", "cuda_runtime.h"), "-c", dst, "-o", obj],
                       check=True, timeout=600)
        objs.append(obj)
    aobj = os.path.join(SCRATCH, "emul_api.o")
    subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-pthread", "-DLJ_EMUL"] + extra + ["-c", os.path.join(HERE, "emul_api.cpp"), "-o", aobj],
                   check=True, timeout=300)
    objs.append(aobj)
    hobj = os.path.join(SCRATCH, "lj_host.o")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off"] + extra + ["-c", os.path.join(PKG, "host", "lj_host.cpp"),
                    "-o", hobj], check=True, timeout=300)
    subprocess.run(["g++", "-shared", "-pthread"] + (["-fsanitize=address"] if asan else []) + ["-o", OUT] + objs + [hobj, "-ldl"],
                   check=True, timeout=300)
    # stand-in NCCL for the slab path: ranks are threads of one process (found through LD_LIBRARY_PATH by lj_nccl.h)
    subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-o", os.path.join(SCRATCH, "libnccl.so.2"),
                    os.path.join(HERE, "nccl_emul.cpp")], check=True, timeout=300) if not asan else None
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, asan="--asan" in sys.argv))
