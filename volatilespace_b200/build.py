"""Builds libvsb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m volatilespace_b200.build            # build if stale
    python -m volatilespace_b200.build --force

The shared library travels to the GPU box with the gpurun snapshot; it is
git-ignored.  nvcc cross-compiles without a GPU.
"""
import os
import shutil
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
    "vsb_gravity_sym_lean.cu": [],
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


def _obj(src):
    return os.path.join(OBJ, src.replace(".cu", "").replace("@", "_") + ".o")


def _read(path):
    try:
        with open(path) as f:
            return f.read()
    except OSError:
        return None


def _write_if_changed(path, text):
    if _read(path) != text:
        with open(path, "w") as f:
            f.write(text)


def status():
    """What the last build() did about the SASS post-pass: {"sym_sched": "resched" | "ptxas",
    "warning": None | reason}.  The same string is compiled into the library (vsb_sym_schedule)."""
    txt = _read(os.path.join(OBJ, "sym_sched.txt")) or "ptxas (never built)"
    return {"sym_sched": txt.split(" ")[0], "warning": txt[txt.index("(") + 1:-1] if "(" in txt else None}


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "vsb_common.cuh"), os.path.join(CSRC, "vsb_orbit_math.cuh"),
               os.path.join(CSRC, "vsb_grid.cuh"), os.path.join(CSRC, "vsb_ddtrig.h"), INCLUDE,
               os.path.abspath(__file__)]
    objs, procs = [], []
    env = dict(os.environ)
    env.pop("CC", None)    # the image exports a gcc wrapper nvcc should not pick up
    env.pop("CXX", None)
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src.split("@")[0])
        o = _obj(src)
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
    # SASS post-pass (sass_resched.py: same instructions, same registers, same results bit for
    # bit -- the accumulates of the pair loop regrouped so that they ride the operand-reuse
    # cache).  ptxas's object is never modified: the patched code goes to a SEPARATE file
    # (<name>_resched.o, rebuilt when ptxas's object or the tool is newer), and which of the two
    # is linked is recorded in _build/sym_sched.txt, compiled into the library
    # (vsb_sym_schedule()) and returned by status().  VSB_NO_RESCHED=1 links ptxas's schedule.
    sched = "resched"
    for src, pattern in RESCHED.items():
        plain = _obj(src)
        patched = plain[:-2] + "_resched.o"
        if os.environ.get("VSB_NO_RESCHED"):
            sched = "ptxas (VSB_NO_RESCHED set at build time)"
            continue
        try:
            tool = os.path.join(HERE, "sass_resched.py")
            if force or _stale(patched, [plain, tool]):
                from volatilespace_b200 import sass_resched
                tmp = patched + ".tmp"
                shutil.copyfile(plain, tmp)
                for line in sass_resched.patch_object(tmp, pattern, no_yield=True):
                    if verbose:
                        print(line)
                os.replace(tmp, patched)
            objs[objs.index(plain)] = patched
        except Exception as e:  # noqa: BLE001 -- ptxas's schedule stays: slower, same results
            sched = "ptxas (SASS post-pass failed: %s)" % str(e).replace("\n", " ")[:160]
            sys.stderr.write("WARNING: sass_resched skipped for %s: %s\n" % (src, e))
    _write_if_changed(os.path.join(OBJ, "sym_sched.txt"), sched)
    info_c = os.path.join(OBJ, "vsb_build_info.c")
    _write_if_changed(info_c, """/* generated by volatilespace_b200/build.py */
#include <stdlib.h>
#include <string.h>
const char *vsb_sym_schedule(void)
{
    const char *e = getenv("VSB_SYM_SCHED");
    if (e && strcmp(e, "ptxas") == 0) return "ptxas (VSB_SYM_SCHED=ptxas)";
    return "%s";
}
""" % sched.replace("\\", "/").replace('"', "'"))
    info_o = os.path.join(OBJ, "vsb_build_info.o")
    if force or _stale(info_o, [info_c]):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-c", info_c, "-o", info_o], env=env)
    objs.append(info_o)
    link_txt = os.path.join(OBJ, "link.txt")
    if force or procs or _stale(LIB, objs) or _read(link_txt) != "\n".join(objs):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd, env=env)
        _write_if_changed(link_txt, "\n".join(objs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
