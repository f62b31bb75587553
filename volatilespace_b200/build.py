"""Builds libvsb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m volatilespace_b200.build            # build if stale
    python -m volatilespace_b200.build --force

The shared library travels to the GPU box with the gpurun snapshot; it is
git-ignored.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libvsb200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include", "vsb200.h")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "--extended-lambda", "-Xcompiler", "-fPIC",
          "-Xcompiler", "-fvisibility=default"]
# per-file flags: the parity-critical files are compiled without FMA contraction
SOURCES = {
    "vsb_api.cu": [],
    "vsb_gravity.cu": [],
    "vsb_gravity_sym.cu": [],
    # second object of the same source: ptxas's own schedule (reference of the SASS post-pass)
    "vsb_gravity_sym.cu@ptxas": ["-DVSB_SYM_PTXAS_SCHEDULE"],
    "vsb_pairs.cu": ["-fmad=false"],
    "vsb_broadphase.cu": ["-fmad=false"],
    "vsb_parents_grid.cu": ["-fmad=false"],
    "vsb_editor.cu": ["-fmad=false"],
    "vsb_kepler.cu": ["-fmad=false"],
    "vsb_visible.cu": ["-fmad=false"],
    "vsb_body.cu": ["-fmad=false"],
    "vsb_convert.cu": ["-fmad=false"],
    "vsb_intersect.cu": ["-fmad=false"],
    "vsb_events.cu": ["-fmad=false"],
    "vsb_segments.cu": ["-fmad=false"],
}


# objects whose pair loop is re-scheduled after ptxas (kernel name pattern)
RESCHED = {"vsb_gravity_sym.cu": "allpairs_sym_kernel"}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "vsb_common.cuh"), os.path.join(CSRC, "vsb_orbit_math.cuh"),
               os.path.join(CSRC, "vsb_grid.cuh"), INCLUDE,
               os.path.abspath(__file__), os.path.join(HERE, "sass_resched.py")]
    objs, procs = [], []
    env = dict(os.environ)
    env.pop("CC", None)    # the image exports a gcc wrapper nvcc should not pick up
    env.pop("CXX", None)
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src.split("@")[0])
        o = os.path.join(OBJ, src.replace(".cu", "").replace("@", "_") + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc()] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE,
                                                stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s\n" % src)
    if failed:
        raise RuntimeError("libvsb200 build failed")
    # SASS post-pass on freshly compiled objects (sass_resched.py: same instructions, same
    # registers, same results bit for bit -- the accumulates of the pair loop regrouped so that
    # they ride the operand-reuse cache).  VSB_NO_RESCHED=1 keeps ptxas's own schedule.
    for src, _ in procs:
        if src in RESCHED and not os.environ.get("VSB_NO_RESCHED"):
            from volatilespace_b200 import sass_resched
            o = os.path.join(OBJ, src.replace(".cu", "").replace("@", "_") + ".o")
            try:
                for line in sass_resched.patch_object(o, RESCHED[src], no_yield=True):
                    if verbose:
                        print(line)
            except Exception as e:  # noqa: BLE001 -- ptxas's schedule stays: slower, same results
                sys.stderr.write("sass_resched skipped for %s: %s\n" % (src, e))
    if force or procs or _stale(LIB, objs):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
