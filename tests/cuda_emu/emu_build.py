"""TEST INFRASTRUCTURE ONLY -- builds tests/cuda_emu/_build/libemu.so: every csrc/*.cu of the product compiled by g++
against the CPU stand-in for the CUDA execution model (cuda_runtime.h in this directory), so that the CPU test suite
can EXECUTE the kernels.  The sources are used as they are, except for three mechanical rewrites that have no C++
spelling:  kernel<<<grid, block, smem, stream>>>(args)  ->  emu::launch(grid, block, smem, [&] { kernel(args); }),
extern __shared__ T name[]  ->  T *name = (T *)emu::dyn_smem,  and inline PTX (a prefetch hint) is dropped."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "lagrangian_microbes_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SOURCES = ["api.cu", "advect.cu", "bin.cu", "strip.cu", "pairs.cu", "interact.cu", "resolve.cu", "analysis.cu", "record.cu"]    # = _lib._SOURCES


def _split_top_level(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


_LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)<<<([^;]*?)>>>\(([^;]*)\);")
_DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\[\];")
_ASM = re.compile(r'asm volatile\("[^"]*"[^;]*;')


def transform(text):
    def launch(m):
        kernel, cfg, args = m.group(1), _split_top_level(m.group(2)), m.group(3)
        assert len(cfg) in (2, 3, 4), cfg
        smem = cfg[2] if len(cfg) > 2 else "0"
        return "emu::launch((unsigned int)(%s), (unsigned int)(%s), (size_t)(%s), [&] { %s(%s); });" % (cfg[0], cfg[1], smem, kernel, args)
    text, n_launch = _LAUNCH.subn(launch, text)
    text = _DYN.sub(lambda m: "%s *%s = reinterpret_cast<%s *>(emu::dyn_smem);" % (m.group(1), m.group(2), m.group(1)), text)
    text = _ASM.sub(";", text)
    assert "<<<" not in text and "extern __shared__" not in text
    return text, n_launch


def _compile_and_link(sources, flags, link_flags, out):
    """One g++ per translation unit, all at once (the CPU suite spends most of its emulator time here), then the link."""
    from concurrent.futures import ThreadPoolExecutor
    objs = [src[:-4] + "." + os.path.basename(out).replace(".", "_") + ".o" for src in sources]
    objs = [os.path.join(OUT, os.path.basename(o)) for o in objs]

    def one(pair):
        src, obj = pair
        subprocess.check_call(["g++", "-c"] + flags + ["-o", obj, src])
    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 4)) as pool:
        list(pool.map(one, zip(sources, objs)))
    subprocess.check_call(["g++"] + link_flags + ["-o", out + ".tmp"] + objs)
    os.replace(out + ".tmp", out)
    return out


def build(force=False):
    so = os.path.join(OUT, "libemu.so")
    deps = [os.path.join(CSRC, f) for f in SOURCES + ["lm_internal.cuh", "philox.cuh"]] + \
           [os.path.join(HERE, f) for f in ("cuda_runtime.h", "emu.cpp", "harness.cpp", "emu_build.py")]
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    os.makedirs(OUT, exist_ok=True)
    cpps = []
    for f in SOURCES:
        text, n = transform(open(os.path.join(CSRC, f)).read())
        assert n > 0 or f == "api.cu", "no kernel launch found in " + f
        path = os.path.join(OUT, f.replace(".cu", "_emu.cpp"))
        with open(path, "w") as fh:
            fh.write(text)
        cpps.append(path)
    flags = ["-std=c++17", "-O1", "-g", "-pthread", "-fPIC", "-ffp-contract=off", "-w", "-I", HERE, "-I", CSRC]
    return _compile_and_link(cpps + [os.path.join(HERE, "emu.cpp"), os.path.join(HERE, "harness.cpp")], flags,
                             ["-shared", "-pthread"], so)


def build_tsan(sanitizer="thread", force=False):
    """tests/cuda_emu/_build/tsan_driver: the emulated library + tsan_driver.cpp under ThreadSanitizer (see the driver);
    sanitizer="address,undefined" builds asan_driver instead: out-of-bounds accesses of global / shared memory."""
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "tsan_driver" if sanitizer == "thread" else "asan_driver")
    deps = [os.path.join(CSRC, f) for f in SOURCES + ["lm_internal.cuh", "philox.cuh"]] + \
           [os.path.join(HERE, f) for f in ("cuda_runtime.h", "emu.cpp", "tsan_driver.cpp", "emu_build.py")]
    if not force and os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return exe
    cpps = []
    for f in SOURCES:
        text, _ = transform(open(os.path.join(CSRC, f)).read())
        path = os.path.join(OUT, f.replace(".cu", "_san_%s.cpp" % sanitizer.split(",")[0]))
        with open(path, "w") as fh:
            fh.write(text)
        cpps.append(path)
    flags = ["-std=c++17", "-O1", "-g", "-pthread", "-fsanitize=" + sanitizer, "-fno-omit-frame-pointer", "-ffp-contract=off", "-w",
             "-I", HERE, "-I", CSRC]
    return _compile_and_link(cpps + [os.path.join(HERE, "emu.cpp"), os.path.join(HERE, "tsan_driver.cpp")], flags,
                             ["-pthread", "-fsanitize=" + sanitizer], exe)


if __name__ == "__main__":
    import sys
    if "--tsan" in sys.argv:
        print(build_tsan(force=True))
    elif "--asan" in sys.argv:
        print(build_tsan("address,undefined", force=True))
    else:
        print(build(force=True))
