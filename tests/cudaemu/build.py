"""Builds tests/cudaemu/_build/libcudaemu.so: csrc/tracegen.cu, csrc/derive.cu and csrc/machine.cpp compiled for the host against the
CPU stand-in of the CUDA runtime (tests/cudaemu/cuda_runtime.h).  The only change to the source text is the launch syntax:
`kernel<<<grid, block, shared, stream>>>(args)` becomes `emu_launch(kernel, grid, block, args)`."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ziren_b200", "csrc")
SOURCES = ("tracegen.cu", "derive.cu", "machine.cpp")


def _split_top_level(s: str) -> list[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def _kernel_name_start(text: str, at: int) -> tuple[int, int]:
    """the kernel expression that ends right before `<<<`: identifier, optionally with template arguments"""
    i = at
    while text[i - 1].isspace():
        i -= 1
    end = i
    if text[i - 1] == ">":
        depth = 0
        while True:
            i -= 1
            if text[i] == ">":
                depth += 1
            elif text[i] == "<":
                depth -= 1
                if depth == 0:
                    break
    while text[i - 1].isalnum() or text[i - 1] in "_:":
        i -= 1
    return i, end


def rewrite_launches(text: str) -> tuple[str, int]:
    """kernel<<<g, b, shm, stream>>>(args) -> emu_launch_dyn((kernel), g, b, shm, args); returns (text, number of launch sites)"""
    n = 0
    while True:
        at = text.find("<<<")
        if at < 0:
            return text, n
        ks, ke = _kernel_name_start(text, at)
        end = text.index(">>>", at)
        cfg = _split_top_level(text[at + 3:end])
        assert len(cfg) in (2, 3, 4), cfg
        k = end + 3
        while text[k].isspace():
            k += 1
        assert text[k] == "(", text[k:k + 20]
        depth, j = 0, k
        while True:
            depth += text[j] == "("
            depth -= text[j] == ")"
            if depth == 0:
                break
            j += 1
        args = text[k + 1:j].strip()
        shm = cfg[2] if len(cfg) > 2 else "0"
        call = f"emu_launch_dyn(({text[ks:ke]}), {cfg[0]}, {cfg[1]}, {shm}" + (", " + args if args else "") + ")"
        text = text[:ks] + call + text[j + 1:]
        n += 1


def poison_shared(text: str) -> tuple[str, int]:
    """`__shared__ T name[n]...;` -> the same static array, filled with 0xDEADBEEF once per block before its first use: shared
    memory is not zero on a GPU and keeps what the previous block left, so a kernel that reads a word it has not written must
    not pass here by reading a stale zero.  `extern __shared__ T name[];` -> a pointer to the launch's dynamic shared memory."""
    text, n_dyn = re.subn(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1* \2 = static_cast<\1*>(emu_dyn_shared());", text)
    text, n = re.subn(r"__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)((?:\[[^\]]*\])*);",
                      r"alignas(16) static \1 \2\3; emu_poison(&\2, sizeof(\2));", text)
    assert "__shared__" not in re.sub(r"//[^\n]*", "", text), "a __shared__ declaration the rewrite does not know"
    return text, n + n_dyn


def build() -> str:
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(HERE, "_build", "libcudaemu.so")
    deps = [os.path.join(HERE, f) for f in ("cuda_runtime.h", "emu_main.cpp", "build.py")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp", ".inc"))]
    if os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    units = [os.path.join(HERE, "emu_main.cpp")]
    sites = 0
    for name in SOURCES:
        text, n = rewrite_launches(open(os.path.join(CSRC, name)).read())
        sites += n
        text, n_shared = poison_shared(text)
        dst = os.path.join(out_dir, os.path.splitext(name)[0] + "_emu.cpp")
        with open(dst, "w") as f:
            f.write(f'#line 1 "{os.path.join(CSRC, name)}"\n' + text)
        units.append(dst)
    assert sites == 11, sites      # 7 in tracegen.cu, 4 in derive.cu
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-I" + HERE, "-I" + CSRC, "-Wno-unknown-pragmas",
                           *units, "-o", so])
    return so


ALL_SOURCES = ("tracegen.cu", "derive.cu", "ntt.cu", "hash.cu", "layout.cu", "logup.cu", "quotient.cu", "open.cu", "fri.cu", "prover.cu",
               "capi.cu", "machine.cpp", "quotient_codegen.cpp")


def replace_ptx(text: str) -> str:
    """the three cp.async statements of csrc/open.cu: a 4-byte copy, and nothing to wait for"""
    text = text.replace('asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");',
                        "*reinterpret_cast<unsigned*>(d) = *reinterpret_cast<const unsigned*>(gmem_src);")
    text = text.replace('asm volatile("cp.async.commit_group;" ::: "memory");', ";")
    text = text.replace('asm volatile("cp.async.wait_group 0;" ::: "memory");', ";")
    assert "asm volatile" not in text
    return text


def build_full() -> str:
    """tests/cudaemu/_build/libzkb200emu.so: EVERY source of libzkb200.so for the host, exporting the same C ABI"""
    out_dir = os.path.join(HERE, "_build", "full")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(HERE, "_build", "libzkb200emu.so")
    deps = [os.path.join(HERE, f) for f in ("cuda_runtime.h", "build.py")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp", ".inc"))]
    if os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    objs = []
    procs = []
    for name in os.listdir(CSRC):          # headers hold device code too (ntt_lean.cuh: dynamic shared memory)
        if name.endswith((".cuh", ".h", ".inc")):
            text, _ = poison_shared(open(os.path.join(CSRC, name)).read())
            with open(os.path.join(out_dir, name), "w") as f:
                f.write(text)
    for name in ALL_SOURCES:
        text, _ = rewrite_launches(open(os.path.join(CSRC, name)).read())
        text, _ = poison_shared(text)
        text = replace_ptx(text).replace('"../../include/zkb200.h"', '"' + os.path.join(ROOT, "include", "zkb200.h") + '"')
        dst = os.path.join(out_dir, os.path.splitext(name)[0] + "_emu.cpp")
        with open(dst, "w") as f:
            f.write(f'#line 1 "{os.path.join(CSRC, name)}"\n' + text)
        obj = dst[:-4] + ".o"
        objs.append(obj)
        procs.append(subprocess.Popen(["g++", "-O2", "-std=c++20", "-fPIC", "-pthread", "-fpermissive", "-w", "-D__CUDACC__", "-I" + HERE, "-I" + out_dir,
                                       "-c", dst, "-o", obj]))
    extra = os.path.join(out_dir, "emu_exports.cpp")
    with open(extra, "w") as f:      # lets the test side mark a host buffer as standing for device / pinned memory
        f.write('#include <cuda_runtime.h>\nextern "C" void emu_register(void* p, size_t n, int type) { emu_allocs.add(p, n, (cudaMemoryType)type); }\n')
    objs.append(extra[:-4] + ".o")
    procs.append(subprocess.Popen(["g++", "-O1", "-std=c++20", "-fPIC", "-I" + HERE, "-c", extra, "-o", objs[-1]]))
    assert all(p.wait() == 0 for p in procs), "emulated build failed"
    subprocess.check_call(["g++", "-shared", "-pthread", *objs, "-ldl", "-o", so])
    return so


if __name__ == "__main__":
    import sys
    if "full" in sys.argv:
        print(build_full())
        sys.exit(0)
    print(build())
