"""TEST INFRASTRUCTURE ONLY -- builds `oracle/_ref/libref_tsdf.so` from the reference's own source and stages the
reference's `back_project.py` (verbatim copy, byte for byte) next to it as `oracle/_ref/back_project.py`, so that
`bench.py` can time the UNMODIFIED reference function on the GPU box (on CUDA: `reference_gpu`; on the host cores:
`cpu_baseline.kind = "reference"` / `--impl reference`), where `/root/reference` does not exist.

The reference's TSDF GPU kernel is a CUDA C string that PyCUDA compiles at run
time (`/root/reference/deep3dmap/core/tsdf/tsdf_volume.py:67-142`).  PyCUDA is
not installed, so this recipe pulls the string out of the reference file *where
it lies* (nothing is copied into the repository: `oracle/_ref/` is git-ignored),
appends a 30-line C launcher that reproduces the reference launch geometry
(`tsdf_volume.py:147-155, 232-256`) and compiles it with nvcc defaults (FMA
contraction ON, exactly what `pycuda.compiler.SourceModule` does).

The resulting shared object travels to the GPU box with the snapshot and is the
"reference GPU kernel" the `-m gpu` parity tests compare against bit-for-bit.
Only `tests/`, `bench.py --impl reference` / `cpu_baseline` and
`__graft_entry__.smoke()` may load it.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = "/root/reference/deep3dmap/core/tsdf/tsdf_volume.py"
REF_BP_FILE = "/root/reference/deep3dmap/core/voxel/back_project.py"
OUT_DIR = os.path.join(HERE, "_ref")

LAUNCHER = r'''
// ---- launcher appended by oracle/build_ref.py (not reference code) ----
#include <cuda_runtime.h>
#include <math.h>
extern "C" int ref_tsdf_integrate(float* tsdf, float* weight, float* color,
                                  int dx, int dy, int dz, const float* origin3,
                                  const float* intr9, const float* pose16,
                                  float voxel_size, int im_h, int im_w, float trunc,
                                  float obs_weight, const float* color_im_dev,
                                  const float* depth_dev, cudaStream_t stream) {
  // launch geometry of tsdf_volume.py:147-155 (MAX_THREADS_PER_BLOCK = 1024)
  const int tpb = 1024;
  double nvox = (double)dx * dy * dz;
  long n_blocks = (long)ceil(nvox / tpb);
  long gx = (long)floor(cbrt((double)n_blocks)); if (gx > 2147483647L) gx = 2147483647L; if (gx < 1) gx = 1;
  long gy = (long)floor(sqrt((double)n_blocks / gx)); if (gy > 65535) gy = 65535; if (gy < 1) gy = 1;
  long gz = (long)ceil((double)n_blocks / (double)(gx * gy)); if (gz > 65535) gz = 65535; if (gz < 1) gz = 1;
  int n_loops = (int)ceil(nvox / ((double)gx * gy * gz * tpb));
  float h_dim[3] = {(float)dx, (float)dy, (float)dz};
  float *d_dim, *d_org, *d_intr, *d_pose, *d_other;
  cudaMalloc(&d_dim, 12); cudaMalloc(&d_org, 12); cudaMalloc(&d_intr, 36);
  cudaMalloc(&d_pose, 64); cudaMalloc(&d_other, 24);
  cudaMemcpyAsync(d_dim, h_dim, 12, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(d_org, origin3, 12, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(d_intr, intr9, 36, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(d_pose, pose16, 64, cudaMemcpyHostToDevice, stream);
  for (int l = 0; l < n_loops; ++l) {
    float h_other[6] = {(float)l, voxel_size, (float)im_h, (float)im_w, trunc, obs_weight};
    cudaMemcpyAsync(d_other, h_other, 24, cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
    integrate<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), tpb, 0, stream>>>(
        tsdf, weight, color, d_dim, d_org, d_intr, d_pose, d_other,
        (float*)color_im_dev, (float*)depth_dev);
  }
  cudaError_t e = cudaStreamSynchronize(stream);
  cudaFree(d_dim); cudaFree(d_org); cudaFree(d_intr); cudaFree(d_pose); cudaFree(d_other);
  return (int)e;
}
'''


def extract_kernel_source():
    src = open(REF_FILE).read()
    m = re.search(r'SourceModule\("""(.*?)"""\)', src, re.S)
    if not m:
        raise RuntimeError("reference CUDA string not found in " + REF_FILE)
    return m.group(1)


def stage_back_project():
    """Byte-for-byte copy of the reference's back_project.py into the git-ignored oracle/_ref/ (travels to the GPU box)."""
    if not os.path.exists(REF_BP_FILE):
        return None
    import shutil
    os.makedirs(OUT_DIR, exist_ok=True)
    dst = os.path.join(OUT_DIR, "back_project.py")
    shutil.copyfile(REF_BP_FILE, dst)
    return dst


def build(verbose=True):
    """Returns the path of the built .so, or None when the reference tree is absent."""
    if not os.path.exists(REF_FILE):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    stage_back_project()
    cu = os.path.join(OUT_DIR, "ref_tsdf_kernel.cu")
    so = os.path.join(OUT_DIR, "libref_tsdf.so")
    with open(cu, "w") as f:
        f.write("// extracted at build time from %s -- NOT tracked by git\n" % REF_FILE)
        f.write(extract_kernel_source())
        f.write(LAUNCHER)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
           "-shared", "-Xcompiler", "-fPIC", "-o", so, cu]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    p = build()
    print("built" if p else "reference tree absent; nothing built", p or "")
    sys.exit(0)
