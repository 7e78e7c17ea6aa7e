// Probe that pins the tcgen05 operand conventions the fused tensor-core kernels rely on (run on a B200):
//   mode 0  D[128 x 64] = A(TMEM, lane = row, column = k) x W[K x 64]        B = MN-major SWIZZLE_128B image of W
//   mode 1  D[128 x Kp] = G(TMEM)[128 x 64] x W^T                            B = the SAME image read K-major
//   mode 2  D[128 x 64] = Aw^T x G,  Aw [nE x 128], G [nE x 64] in smem      both operands MN-major SWIZZLE_128B
//   mode 3  dW[64 x 64] = Z^T x G,  Z [nE x 64], G [nE x 64] in smem         M = 64, full 3xTF32; prints the TMEM lane of each row
// Each mode is checked against a double-precision CPU product; passes = 1 (plain TF32) or 3 (hi/lo split, ~fp32).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu && ./umma_probe <mode> [passes] [Kp]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../neuralgraphpde.jl_b200/csrc/ngpde_umma.cuh"

using namespace ngpde::umma;

struct Params {
  int mode, passes, K, Kp, nE, img;  // img: 0 = SWIZZLE_128B (16-byte base), 1 = SWIZZLE_128B_BASE32B, 2 = no swizzle
};

// float offset of element (r, c) -- r = row of 128 bytes (k for a weight image, edge for staging), c in [0, 32*groups)
__device__ __forceinline__ uint32_t img_offset(int img, int rows, int ncols, int r, int c) {
  if (img == 0) return sw128_offset(c >> 5, rows, r, c & 31);
  if (img == 1 || img == 3) return sw128b32_offset(c >> 5, rows, r, c & 31);
  return (uint32_t)(((r >> 3) * (ncols >> 2) + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3));  // core matrices [r/8][c/4]
}
// descriptor of the image read with its 128-byte rows as the operand's K index ("MN-major"), at K-step ks (8 rows)
__device__ __forceinline__ uint64_t desc_rows_are_k(int img, uint32_t base, int rows, int ncols, int ks) {
  if (img == 0) return make_sdesc(base + ks * 1024, rows * 128, 1024, 2);
  if (img == 1) return make_sdesc(base + ks * 1024, rows * 128, 512, 1);
  return make_sdesc(base + ks * (ncols >> 2) * 128, /*lbo: k-groups*/ (ncols >> 2) * 128, /*sbo: MN chunks*/ 128, 0);
}
// descriptor of the image read with its rows as the operand's N index ("K-major"), at K-step ks (8 floats of the row)
__device__ __forceinline__ uint64_t desc_rows_are_n(int img, uint32_t base, int rows, int ncols, int ks) {
  if (img == 0) return make_sdesc(base + (ks >> 2) * (rows * 128) + (ks & 3) * 32, 0, 1024, 2);
  if (img == 1) return make_sdesc(base + (ks >> 2) * (rows * 128) + (ks & 3) * 32, 0, 1024, 1);  // experiment
  if (img == 3) return make_sdesc(base + (ks >> 2) * (rows * 128) + (ks & 3) * 32, 0, 512, 1);   // experiment
  return make_sdesc(base + ks * 2 * 128, /*lbo: K chunks*/ 128, /*sbo: 8-row groups*/ (ncols >> 2) * 128, 0);
}

__global__ void __launch_bounds__(128) probe_kernel(Params p, const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t COL_D = 0, COL_AHI = 128, COL_ALO = 256;

  float* Bhi = reinterpret_cast<float*>(smem);                 // operand images
  float* Blo = reinterpret_cast<float*>(smem + 64 * 1024);
  float* Aw = reinterpret_cast<float*>(smem + 128 * 1024);     // mode 2: A operand [4 groups][nE][32]

  if (p.mode == 0 || p.mode == 1) {
    // W image: [half][k][32 floats], 16-byte chunks XOR-swizzled with (k & 7)
    for (int i = tid; i < p.Kp * 64; i += 128) {
      const int k = i / 64, n = i % 64;
      const float w = (k < p.K) ? B[k * 64 + n] : 0.f;
      const float hi = tf32_hi(w);
      const uint32_t off = img_offset(p.img, p.Kp, 64, k, n);
      Bhi[off] = hi;
      Blo[off] = tf32_hi(w - hi);
    }
    // A rows into TMEM: mode 0 has K columns (padded to Kp), mode 1 has 64
    const int ka = p.mode == 0 ? p.Kp : 64, kv = p.mode == 0 ? p.K : 64;
    for (int c0 = 0; c0 < ka; c0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = (c0 + j < kv) ? A[tid * kv + c0 + j] : 0.f;
        const float h = tf32_hi(a);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(tf32_hi(a - h));
      }
      tmem_st8(lane_base + COL_AHI + c0, hi);
      tmem_st8(lane_base + COL_ALO + c0, lo);
    }
    tmem_wait_st();
  } else if (p.mode == 3) {
    float* Alo = reinterpret_cast<float*>(smem + 160 * 1024);
    for (int i = tid; i < p.nE * 64; i += 128) {
      const int e = i / 64, m = i % 64;
      const float a = A[e * 64 + m], g = B[e * 64 + m];
      const float ah = tf32_hi(a), gh = tf32_hi(g);
      const uint32_t off = img_offset(p.img, p.nE, 64, e, m);
      Aw[off] = ah;
      Alo[off] = tf32_hi(a - ah);
      Bhi[off] = gh;
      Blo[off] = tf32_hi(g - gh);
    }
    // sentinel in the accumulator columns so that untouched lanes are recognisable
    for (int c0 = 0; c0 < 64; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(-777.f);
      tmem_st8(lane_base + COL_D + c0, v);
    }
    tmem_wait_st();
  } else {
    // mode 2: Aw [nE][128] -> 4 groups, G [nE][64] -> 2 groups (hi only unless passes == 3)
    for (int i = tid; i < p.nE * 128; i += 128) {
      const int e = i / 128, m = i % 128;
      Aw[img_offset(p.img, p.nE, 128, e, m)] = tf32_hi(A[e * 128 + m]);
    }
    for (int i = tid; i < p.nE * 64; i += 128) {
      const int e = i / 64, n = i % 64;
      const float g = B[e * 64 + n];
      const float hi = tf32_hi(g);
      const uint32_t off = img_offset(p.img, p.nE, 64, e, n);
      Bhi[off] = hi;
      Blo[off] = tf32_hi(g - hi);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    const uint32_t bhi = smem_u32(Bhi), blo = smem_u32(Blo);
    if (p.mode == 0) {
      const uint32_t idesc = make_idesc(128, 64, /*a_mn=*/0, /*b_mn=*/1);
      int first = 1;
      for (int pass = 0; pass < p.passes; ++pass) {
        const uint32_t acol = (pass == 1) ? COL_ALO : COL_AHI;
        const uint32_t bb = (pass == 2) ? blo : bhi;
        for (int ks = 0; ks < p.Kp / 8; ++ks) {
          const uint64_t bdesc = desc_rows_are_k(p.img, bb, p.Kp, 64, ks);
          mma_tf32_ts(tmem + COL_D, tmem + acol + ks * 8, bdesc, idesc, !first);
          first = 0;
        }
      }
    } else if (p.mode == 1) {
      const uint32_t idesc = make_idesc(128, p.Kp, 0, /*b_mn=*/0);
      int first = 1;
      for (int pass = 0; pass < p.passes; ++pass) {
        const uint32_t acol = (pass == 1) ? COL_ALO : COL_AHI;
        const uint32_t bb = (pass == 2) ? blo : bhi;
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bdesc = desc_rows_are_n(p.img, bb, p.Kp, 64, ks);
          mma_tf32_ts(tmem + COL_D, tmem + acol + ks * 8, bdesc, idesc, !first);
          first = 0;
        }
      }
    } else if (p.mode == 3) {
      const uint32_t idesc = make_idesc(64, 64, 1, 1);
      const uint32_t ahi = smem_u32(Aw), alo = ahi + 32 * 1024;
      int first = 1;
      for (int pass = 0; pass < p.passes; ++pass) {
        const uint32_t aa = (pass == 1) ? alo : ahi;
        const uint32_t bb = (pass == 2) ? blo : bhi;
        for (int ks = 0; ks < p.nE / 8; ++ks) {
          const uint64_t adesc = desc_rows_are_k(p.img, aa, p.nE, 64, ks);
          const uint64_t bdesc = desc_rows_are_k(p.img, bb, p.nE, 64, ks);
          mma_tf32_ss(tmem + COL_D, adesc, bdesc, idesc, !first);
          first = 0;
        }
      }
    } else {
      const uint32_t idesc = make_idesc(128, 64, 1, 1);
      const uint32_t aw = smem_u32(Aw);
      int first = 1;
      for (int pass = 0; pass < (p.passes == 3 ? 2 : 1); ++pass) {
        const uint32_t bb = pass ? blo : bhi;
        for (int ks = 0; ks < p.nE / 8; ++ks) {
          const uint64_t adesc = desc_rows_are_k(p.img, aw, p.nE, 128, ks);
          const uint64_t bdesc = desc_rows_are_k(p.img, bb, p.nE, 64, ks);
          mma_tf32_ss(tmem + COL_D, adesc, bdesc, idesc, !first);
          first = 0;
        }
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int ncol = p.mode == 1 ? p.Kp : 64;
  for (int c0 = 0; c0 < ncol; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(lane_base + COL_D + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) D[tid * ncol + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float tf32_trunc_host(float x) { return tf32_hi(x); }  // the same rounding the device split uses

int main(int argc, char** argv) {
  Params p{};
  p.mode = argc > 1 ? atoi(argv[1]) : 0;
  p.passes = argc > 2 ? atoi(argv[2]) : 1;
  p.K = argc > 3 ? atoi(argv[3]) : 64;
  p.Kp = (p.K + 15) / 16 * 16;
  p.nE = argc > 4 ? atoi(argv[4]) : 128;
  p.img = argc > 5 ? atoi(argv[5]) : 0;
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  int ar, ac, br, bc, dr = 128, dc;
  if (p.mode == 0) { ar = 128; ac = p.K; br = p.K; bc = 64; dc = 64; }
  else if (p.mode == 1) { ar = 128; ac = 64; br = p.K; bc = 64; dc = p.Kp; }
  else if (p.mode == 3) { ar = p.nE; ac = 64; br = p.nE; bc = 64; dc = 64; }
  else { ar = p.nE; ac = 128; br = p.nE; bc = 64; dc = 64; }
  std::vector<float> A(ar * ac), B(br * bc), D(dr * dc, -777.f);
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 225 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(p, dA, dB, dD);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("mode %d: CUDA error %s\n", p.mode, cudaGetErrorString(err)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  if (p.mode == 3) {
    // find, for every row k of dW = Z^T G, the TMEM lane that holds it
    double maxref = 0, maxerr = 0;
    int lane_of[64], bad = 0;
    for (int k = 0; k < 64; ++k) {
      double best = 1e30; int bl = -1;
      std::vector<double> ref(64, 0.0);
      for (int n = 0; n < 64; ++n) {
        for (int e = 0; e < p.nE; ++e) {
          const double a = p.passes == 1 ? tf32_trunc_host(A[e * 64 + k]) : A[e * 64 + k];
          const double g = p.passes == 1 ? tf32_trunc_host(B[e * 64 + n]) : B[e * 64 + n];
          ref[n] += a * g;
        }
        maxref = fmax(maxref, fabs(ref[n]));
      }
      for (int l = 0; l < 128; ++l) {
        double err = 0;
        for (int n = 0; n < 64; ++n) err = fmax(err, fabs(ref[n] - D[l * 64 + n]));
        if (err < best) { best = err; bl = l; }
      }
      lane_of[k] = bl;
      maxerr = fmax(maxerr, best);
    }
    int touched = 0;
    for (int l = 0; l < 128; ++l) touched += D[l * 64] != -777.f;
    printf("mode 3 img %d passes %d nE %d: max|err| %.3e max|ref| %.3e rel %.3e touched lanes %d; lane of row k:", p.img, p.passes,
           p.nE, maxerr, maxref, maxerr / maxref, touched);
    for (int k = 0; k < 64; ++k) printf(" %d", lane_of[k]);
    printf("\n");
    (void)bad;
    return 0;
  }
  double maxref = 0, maxerr = 0;
  for (int m = 0; m < dr; ++m)
    for (int n = 0; n < dc; ++n) {
      double ref = 0;
      if (p.mode == 0) for (int k = 0; k < p.K; ++k) ref += (double)A[m * ac + k] * B[k * 64 + n];
      else if (p.mode == 1) { if (n < p.K) for (int j = 0; j < 64; ++j) ref += (double)A[m * 64 + j] * B[n * 64 + j]; }
      else for (int e = 0; e < p.nE; ++e) ref += (double)tf32_trunc_host(A[e * 128 + m]) * B[e * 64 + n];
      maxref = fmax(maxref, fabs(ref));
      maxerr = fmax(maxerr, fabs(ref - D[m * dc + n]));
    }
  printf("mode %d img %d passes %d K %d Kp %d nE %d: max|err| %.3e  max|ref| %.3e  rel %.3e  D[0][0..3] = %g %g %g %g\n", p.mode, p.img,
         p.passes, p.K, p.Kp, p.nE, maxerr, maxref, maxerr / maxref, D[0], D[1], D[2], D[3]);
  return 0;
}
