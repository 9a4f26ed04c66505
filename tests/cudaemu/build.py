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


def rewrite_launches(text: str) -> tuple[str, int]:
    """kernel<<<g, b, shm, stream>>>(args) -> emu_launch(kernel, g, b, args); returns (text, number of launch sites)"""
    n = 0
    while True:
        at = text.find("<<<")
        if at < 0:
            return text, n
        m = re.search(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*$", text[:at])
        assert m, "no kernel name before <<<"
        end = text.index(">>>", at)
        cfg = _split_top_level(text[at + 3:end])
        assert len(cfg) in (2, 3, 4), cfg
        assert len(cfg) < 3 or cfg[2] == "0", "dynamic shared memory is not emulated"
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
        call = f"emu_launch({m.group(1)}, {cfg[0]}, {cfg[1]}" + (", " + args if args else "") + ")"
        text = text[:m.start(1)] + call + text[j + 1:]
        n += 1


def poison_shared(text: str) -> tuple[str, int]:
    """`__shared__ T name[n];` -> the same static array, filled with 0xDEADBEEF once per block before its first use: shared
    memory is not zero on a GPU and keeps what the previous block left, so a kernel that reads a word it has not written must
    not pass here by reading a stale zero."""
    return re.subn(r"__shared__\s+(\w+)\s+(\w+)\[([^\]]*)\];", r"static \1 \2[\3]; emu_poison(\2, sizeof(\2));", text)


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
        assert n_shared == (2 if name == "tracegen.cu" else 0), (name, n_shared)
        dst = os.path.join(out_dir, os.path.splitext(name)[0] + "_emu.cpp")
        with open(dst, "w") as f:
            f.write(f'#line 1 "{os.path.join(CSRC, name)}"\n' + text)
        units.append(dst)
    assert sites == 11, sites      # 7 in tracegen.cu, 4 in derive.cu
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-I" + HERE, "-I" + CSRC, "-Wno-unknown-pragmas",
                           *units, "-o", so])
    return so


if __name__ == "__main__":
    print(build())
