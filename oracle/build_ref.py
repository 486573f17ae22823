"""Recipe that compiles the REFERENCE's own CPU fused-MoE translation units, from where they lie under
/root/reference, into oracle/_ref/libref_moe.so (git-ignored; it travels to the GPU box with the snapshot).

    python oracle/build_ref.py            # needs /root/reference (build container only)

Direct g++ on six reference sources (csrc/cpu/cpu_fused_moe.cpp, mla_decode.cpp, layernorm.cpp, pos_encoding.cpp,
activation.cpp and utils.cpp with the reference's own VLLM_NUMA_DISABLED switch) plus oracle/ref_shim.cpp (ours); no cmake,
no reference build system.  The only external
dependency is the PyTorch C++ headers / libraries of this image (the reference's CPU kernels take torch tensors).
TEST INFRASTRUCTURE / CPU BASELINE ONLY.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LVLLM_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref", "libref_moe.so")
ISA_FLAGS = ["-mf16c", "-mfma", "-mavx2", "-mavx512f", "-mavx512bw", "-mavx512vl", "-mavx512dq", "-mavx512bf16", "-mavx512vnni", "-mamx-tile",
             "-mamx-bf16", "-mamx-int8"]


def build(force: bool = False) -> str | None:
    srcs = [os.path.join(REF, "csrc", "cpu", f) for f in ("cpu_fused_moe.cpp", "mla_decode.cpp", "utils.cpp", "layernorm.cpp",
                                                          "pos_encoding.cpp", "activation.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        return OUT if os.path.exists(OUT) else None          # GPU box: only the prebuilt file exists
    shim = os.path.join(HERE, "ref_shim.cpp")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(s) for s in srcs + [shim]):
        return OUT
    import torch
    tdir = os.path.dirname(torch.__file__)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    import shutil
    import tempfile
    objdir = tempfile.mkdtemp(prefix="ref_obj_")   # objects do not travel with the snapshot: only the library does
    cflags = ["-O3", "-std=c++17", "-fopenmp", "-fPIC", "-DVLLM_NUMA_DISABLED", "-D_GLIBCXX_USE_CXX11_ABI=1", *ISA_FLAGS,
              "-I", os.path.join(REF, "csrc"), "-I", os.path.join(tdir, "include"),
              "-I", os.path.join(tdir, "include", "torch", "csrc", "api", "include"), "-I", sysconfig.get_paths()["include"]]
    # one g++ per translation unit, all at once (the reference's kernels are template-heavy: ~2.5 min serially), then a link
    procs, objs = [], []
    for src in srcs + [shim]:
        obj = os.path.join(objdir, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen(["g++", *cflags, "-c", src, "-o", obj], stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out[-4000:])
            raise RuntimeError(f"reference CPU kernels did not compile ({os.path.basename(src)})")
    cmd = ["g++", "-shared", "-fopenmp", *objs, "-L", os.path.join(tdir, "lib"), "-lc10", "-ltorch", "-ltorch_cpu",
           "-Wl,-rpath," + os.path.join(tdir, "lib"), "-o", OUT]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    shutil.rmtree(objdir, ignore_errors=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:])
        raise RuntimeError("reference CPU kernels did not link")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
