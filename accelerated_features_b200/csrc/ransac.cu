// Geometric verification of the matches, on the GPU, for a whole batch of pairs at once.  Every caller of the
// reference runs a robust estimator right after the path: cv2.findHomography(pts1, pts2, cv2.USAC_MAGSAC, thr, maxIters=700,
// confidence=0.995) in realtime_demo.py:225 and the notebooks, poselib.estimate_relative_pose in the 1500-pair benchmarks
// (modules/eval/megadepth1500.py:98-113).  Both algorithms live in un-vendored third-party packages (opencv-contrib 4.10.0.84
// is pinned in requirements.txt; poselib is not pinned), so what is restated here is the published scheme they share:
//   hypothesise from minimal samples -> score every hypothesis on all correspondences -> keep the best -> re-fit on its
//   inliers (local optimisation) -> classify with the threshold,
// with MSAC scoring (truncated squared error).  Parity is anchored on the call sites: same inputs, same threshold meaning
// (pixels of the second image for the homography, Sampson distance in normalised image coordinates for the essential matrix);
// tests compare inlier sets and model error with cv2 on the same correspondences (stochastic: agreement, not bit equality).
//
// One launch scores ALL hypotheses of ALL pairs: a block caches the (Hartley-normalised) correspondences of its pair in
// shared memory; a warp owns a hypothesis: lane 0 solves the minimal problem (fp64, a few hundred flops), the 32 lanes score
// it.  A second launch per pair selects the best hypothesis, re-fits by least squares on the inliers (normal equations in
// fp64, smallest eigenvector by cyclic Jacobi) twice, and writes model, mask and count.  Nothing returns to the host.
#include "common.cuh"

namespace xf {

constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32;
constexpr int RS_E_REFITS = 8;         // local-optimisation rounds of the essential-matrix winner (ransac_e_final_kernel)
constexpr int RS_MAX_PTS = 8192;       // correspondences per pair cached in shared memory (4 floats each)

__device__ __forceinline__ uint32_t rs_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// Hartley normalisation of one point set: centroid to the origin, mean distance sqrt(2).  Returns (cx, cy, s).
struct RsNorm { float cx, cy, s; };

__device__ RsNorm rs_normaliser(const float* __restrict__ p, int n, float* red /* >= 3 * RS_WARPS floats */) {
  float sx = 0.f, sy = 0.f;
  for (int i = threadIdx.x; i < n; i += RS_THREADS) { sx += p[2 * i]; sy += p[2 * i + 1]; }
  for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[w] = sx; red[RS_WARPS + w] = sy; }
  __syncthreads();
  float cx = 0.f, cy = 0.f;
  for (int i = 0; i < RS_WARPS; ++i) { cx += red[i]; cy += red[RS_WARPS + i]; }
  cx /= (float)max(n, 1); cy /= (float)max(n, 1);
  __syncthreads();
  float sd = 0.f;
  for (int i = threadIdx.x; i < n; i += RS_THREADS) { const float dx = p[2 * i] - cx, dy = p[2 * i + 1] - cy; sd += sqrtf(dx * dx + dy * dy); }
  for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
  if (l == 0) red[2 * RS_WARPS + w] = sd;
  __syncthreads();
  float md = 0.f;
  for (int i = 0; i < RS_WARPS; ++i) md += red[2 * RS_WARPS + i];
  md /= (float)max(n, 1);
  __syncthreads();
  RsNorm r;
  r.cx = cx; r.cy = cy; r.s = (md > 1e-12f) ? 1.41421356f / md : 1.f;
  return r;
}

// Solve the 8x8 system of a 4-point homography (h22 = 1) by Gaussian elimination with partial pivoting (fp64).
__device__ bool rs_solve_h4(const float (&x)[4], const float (&y)[4], const float (&u)[4], const float (&v)[4], double (&h)[9]) {
  double A[8][9];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double X = x[i], Y = y[i], U = u[i], V = v[i];
    double* r0 = A[2 * i]; double* r1 = A[2 * i + 1];
    r0[0] = X; r0[1] = Y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -U * X; r0[7] = -U * Y; r0[8] = U;
    r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = X; r1[4] = Y; r1[5] = 1; r1[6] = -V * X; r1[7] = -V * Y; r1[8] = V;
  }
  for (int c = 0; c < 8; ++c) {
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < 8; ++r)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (best < 1e-10) return false;                       // degenerate sample (three points on a line, repeated point)
    if (piv != c)
      for (int k = c; k < 9; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1; r < 8; ++r) {
      const double f = A[r][c] * inv;
      for (int k = c; k < 9; ++k) A[r][k] -= f * A[c][k];
    }
  }
  for (int c = 7; c >= 0; --c) {
    double s = A[c][8];
    for (int k = c + 1; k < 8; ++k) s -= A[c][k] * h[k];
    h[c] = s / A[c][c];
  }
  h[8] = 1.0;
  return true;
}

// Smallest-eigenvalue eigenvector of a symmetric N x N matrix (cyclic Jacobi, fp64); M is destroyed.
template <int N>
__device__ void rs_smallest_eigvec(double (&M)[N][N], double (&vec)[N]) {
  double V[N][N];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) off += M[i][j] * M[i][j];
    if (off < 1e-30) break;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) {
        if (fabs(M[p][q]) < 1e-300) continue;
        const double theta = (M[q][q] - M[p][p]) / (2.0 * M[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {
          const double mkp = M[k][p], mkq = M[k][q];
          M[k][p] = c * mkp - s * mkq;
          M[k][q] = s * mkp + c * mkq;
        }
        for (int k = 0; k < N; ++k) {
          const double mpk = M[p][k], mqk = M[q][k];
          M[p][k] = c * mpk - s * mqk;
          M[q][k] = s * mpk + c * mqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < N; ++i)
    if (M[i][i] < M[best][best]) best = i;
  for (int k = 0; k < N; ++k) vec[k] = V[k][best];
}

// squared forward transfer error of a homography (normalised coordinates); invalid (point at infinity) -> huge
__device__ __forceinline__ float rs_h_err2(const float (&h)[9], float x, float y, float u, float v) {
  const float w = h[6] * x + h[7] * y + h[8];
  if (fabsf(w) < 1e-8f) return 1e30f;
  const float iw = 1.0f / w;
  const float du = (h[0] * x + h[1] * y + h[2]) * iw - u, dv = (h[3] * x + h[4] * y + h[5]) * iw - v;
  return du * du + dv * dv;
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 1: hypotheses.  grid (blocks_per_pair, batch); every warp scores `hyp_per_warp` hypotheses and keeps its best.
//   ws_best: [batch][blocks_per_pair] (cost, 9 model floats, normalisers) records
// ---------------------------------------------------------------------------------------------------------------------
struct RsRecord {
  float cost;
  float m[9];
};

__global__ void __launch_bounds__(RS_THREADS) ransac_h_hyp_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                                                  const int* __restrict__ np, int n_max, float thr,
                                                                  int hyp_per_warp, uint32_t seed, RsRecord* __restrict__ best) {
  extern __shared__ float4 sPts[];                 // (x, y, u, v) normalised
  __shared__ float red[3 * RS_WARPS];
  __shared__ RsRecord sBest[RS_WARPS];
  const int pair = blockIdx.y;
  const int n = min(np ? np[pair] : n_max, min(n_max, RS_MAX_PTS));
  RsRecord mine;
  mine.cost = 3.0e38f;
#pragma unroll
  for (int k = 0; k < 9; ++k) mine.m[k] = (k % 4 == 0) ? 1.f : 0.f;
  if (n >= 4) {
    const float* q0 = p0 + (int64_t)pair * n_max * 2;
    const float* q1 = p1 + (int64_t)pair * n_max * 2;
    const RsNorm n0 = rs_normaliser(q0, n, red), n1 = rs_normaliser(q1, n, red);
    for (int i = threadIdx.x; i < n; i += RS_THREADS)
      sPts[i] = make_float4((q0[2 * i] - n0.cx) * n0.s, (q0[2 * i + 1] - n0.cy) * n0.s, (q1[2 * i] - n1.cx) * n1.s,
                            (q1[2 * i + 1] - n1.cy) * n1.s);
    __syncthreads();
    const float t2 = (thr * n1.s) * (thr * n1.s);  // pixel threshold of image 1 in its normalised frame
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int it = 0; it < hyp_per_warp; ++it) {
      const uint32_t hid = (uint32_t)((blockIdx.x * RS_WARPS + warp) * hyp_per_warp + it);
      float h[9];
      int ok = 0;
      if (lane == 0) {
        int idx[4];
        uint32_t s = rs_hash(seed ^ (0x9e3779b9u * (uint32_t)(pair + 1)) ^ (0x85ebca6bu * (hid + 1)));
        for (int k = 0; k < 4; ++k) {
          bool dup;
          do {
            s = rs_hash(s + 0x6d2b79f5u);
            idx[k] = (int)(((uint64_t)s * (uint64_t)n) >> 32);
            dup = false;
            for (int j = 0; j < k; ++j) dup |= (idx[j] == idx[k]);
          } while (dup);
        }
        float x[4], y[4], u[4], v[4];
        for (int k = 0; k < 4; ++k) { const float4 c = sPts[idx[k]]; x[k] = c.x; y[k] = c.y; u[k] = c.z; v[k] = c.w; }
        double hd[9];
        ok = rs_solve_h4(x, y, u, v, hd) ? 1 : 0;
        if (ok)
          for (int k = 0; k < 9; ++k) h[k] = (float)hd[k];
      }
      ok = __shfl_sync(0xffffffffu, ok, 0);
      if (!ok) continue;
#pragma unroll
      for (int k = 0; k < 9; ++k) h[k] = __shfl_sync(0xffffffffu, h[k], 0);
      float cost = 0.f;
      for (int i = lane; i < n; i += 32) {
        const float4 c = sPts[i];
        cost += fminf(rs_h_err2(h, c.x, c.y, c.z, c.w), t2);      // MSAC
      }
      for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
      if (cost < mine.cost) {
        mine.cost = cost;
#pragma unroll
        for (int k = 0; k < 9; ++k) mine.m[k] = h[k];
      }
    }
    if (lane == 0) sBest[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      RsRecord b = sBest[0];
      for (int w = 1; w < RS_WARPS; ++w)
        if (sBest[w].cost < b.cost) b = sBest[w];
      best[(int64_t)pair * gridDim.x + blockIdx.x] = b;
    }
  } else if (threadIdx.x == 0) {
    best[(int64_t)pair * gridDim.x + blockIdx.x] = mine;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 2: one block per pair: best hypothesis -> two rounds of least-squares re-fit on its inliers -> model, mask, count
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RS_THREADS) ransac_h_final_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                                                    const int* __restrict__ np, int n_max, float thr,
                                                                    const RsRecord* __restrict__ best, int blocks_per_pair,
                                                                    float* __restrict__ H_out, unsigned char* __restrict__ mask,
                                                                    int* __restrict__ n_inl) {
  extern __shared__ float4 sPts[];
  __shared__ float red[3 * RS_WARPS];
  __shared__ double sAcc[RS_WARPS][45];
  __shared__ float sH[9];
  __shared__ int sCnt;
  const int pair = blockIdx.x;
  const int n_all = np ? np[pair] : n_max;
  const int n = min(n_all, min(n_max, RS_MAX_PTS));
  unsigned char* mk = mask + (int64_t)pair * n_max;
  for (int i = threadIdx.x; i < n_max; i += RS_THREADS) mk[i] = 0;
  if (n < 4) {
    if (threadIdx.x == 0) {
      n_inl[pair] = n_all < 0 ? n_all : 0;          // a negative count (XF_N_OVERFLOW) propagates
      for (int k = 0; k < 9; ++k) H_out[pair * 9 + k] = (k % 4 == 0) ? 1.f : 0.f;
    }
    return;
  }
  const float* q0 = p0 + (int64_t)pair * n_max * 2;
  const float* q1 = p1 + (int64_t)pair * n_max * 2;
  const RsNorm n0 = rs_normaliser(q0, n, red), n1 = rs_normaliser(q1, n, red);
  for (int i = threadIdx.x; i < n; i += RS_THREADS)
    sPts[i] = make_float4((q0[2 * i] - n0.cx) * n0.s, (q0[2 * i + 1] - n0.cy) * n0.s, (q1[2 * i] - n1.cx) * n1.s,
                          (q1[2 * i + 1] - n1.cy) * n1.s);
  if (threadIdx.x == 0) {
    RsRecord b = best[(int64_t)pair * blocks_per_pair];
    for (int k = 1; k < blocks_per_pair; ++k) {
      const RsRecord c = best[(int64_t)pair * blocks_per_pair + k];
      if (c.cost < b.cost) b = c;
    }
    for (int k = 0; k < 9; ++k) sH[k] = b.m[k];
  }
  __syncthreads();
  const float t2 = (thr * n1.s) * (thr * n1.s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int round = 0; round < 2; ++round) {
    // normal equations of the DLT rows of the current inliers: A^T A (9 x 9 symmetric, 45 entries)
    float h[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) h[k] = sH[k];
    double acc[45];
#pragma unroll
    for (int k = 0; k < 45; ++k) acc[k] = 0.0;
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += RS_THREADS) {
      const float4 c = sPts[i];
      if (rs_h_err2(h, c.x, c.y, c.z, c.w) < t2) {
        ++cnt;
        const double r0[9] = {c.x, c.y, 1.0, 0, 0, 0, -(double)c.z * c.x, -(double)c.z * c.y, -(double)c.z};
        const double r1[9] = {0, 0, 0, c.x, c.y, 1.0, -(double)c.w * c.x, -(double)c.w * c.y, -(double)c.w};
        int e = 0;
#pragma unroll
        for (int a = 0; a < 9; ++a)
#pragma unroll
          for (int b2 = a; b2 < 9; ++b2) acc[e++] += r0[a] * r0[b2] + r1[a] * r1[b2];
      }
    }
#pragma unroll
    for (int k = 0; k < 45; ++k) {
      double v = acc[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sAcc[warp][k] = v;
    }
    cnt = __syncthreads_count(0) * 0 + cnt;      // (barrier; the count is reduced below)
    __syncthreads();
    int total = 0;
    {
      int c2 = cnt;
      for (int o = 16; o > 0; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
      if (lane == 0) red[warp] = (float)c2;
      __syncthreads();
      for (int w = 0; w < RS_WARPS; ++w) total += (int)red[w];
      __syncthreads();
    }
    if (threadIdx.x == 0 && total >= 4) {
      double M[9][9], vec[9];
      int e = 0;
      for (int a = 0; a < 9; ++a)
        for (int b2 = a; b2 < 9; ++b2) {
          double v = 0.0;
          for (int w = 0; w < RS_WARPS; ++w) v += sAcc[w][e];
          M[a][b2] = v; M[b2][a] = v;
          ++e;
        }
      rs_smallest_eigvec<9>(M, vec);
      if (fabs(vec[8]) > 1e-12)
        for (int k = 0; k < 9; ++k) sH[k] = (float)(vec[k] / vec[8]);
    }
    __syncthreads();
  }
  // final classification + de-normalisation  H = T1^-1 Hn T0
  float h[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) h[k] = sH[k];
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += RS_THREADS) {
    const float4 c = sPts[i];
    const bool in = rs_h_err2(h, c.x, c.y, c.z, c.w) < t2;
    mk[i] = in ? 1 : 0;
    cnt += in ? 1 : 0;
  }
  if (threadIdx.x == 0) sCnt = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) atomicAdd(&sCnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    n_inl[pair] = sCnt;
    // T0 = [s0 0 -s0 cx0; 0 s0 -s0 cy0; 0 0 1],  T1^-1 = [1/s1 0 cx1; 0 1/s1 cy1; 0 0 1]
    double Hn[3][3], T0[3][3] = {{n0.s, 0, -n0.s * n0.cx}, {0, n0.s, -n0.s * n0.cy}, {0, 0, 1}};
    double T1i[3][3] = {{1.0 / n1.s, 0, n1.cx}, {0, 1.0 / n1.s, n1.cy}, {0, 0, 1}};
    for (int a = 0; a < 3; ++a)
      for (int b2 = 0; b2 < 3; ++b2) Hn[a][b2] = h[3 * a + b2];
    double tmp[3][3], Hf[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b2 = 0; b2 < 3; ++b2) { double s = 0; for (int k = 0; k < 3; ++k) s += Hn[a][k] * T0[k][b2]; tmp[a][b2] = s; }
    for (int a = 0; a < 3; ++a)
      for (int b2 = 0; b2 < 3; ++b2) { double s = 0; for (int k = 0; k < 3; ++k) s += T1i[a][k] * tmp[k][b2]; Hf[a][b2] = s; }
    const double sc = fabs(Hf[2][2]) > 1e-12 ? 1.0 / Hf[2][2] : 1.0;
    for (int a = 0; a < 3; ++a)
      for (int b2 = 0; b2 < 3; ++b2) H_out[pair * 9 + 3 * a + b2] = (float)(Hf[a][b2] * sc);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Essential matrix (calibrated relative pose; the estimator the 1500-pair benchmarks call through poselib,
// modules/eval/megadepth1500.py:98-113).  Inputs are NORMALISED image coordinates (K^-1 applied by the caller); the minimal
// solver is the normalised 8-point algorithm projected onto the essential manifold (singular values (1,1,0)), scored with the
// Sampson distance.  (poselib uses the 5-point solver, which also handles planar scenes; the 8-point form does not -- see
// DESIGN.md.)  Every LANE solves its own hypothesis (a 9x9 and a 3x3 symmetric eigen-problem in fp64), then the warp scores the
// 32 hypotheses one after the other.
// ---------------------------------------------------------------------------------------------------------------------
__device__ void rs_project_essential(double (&E)[9]) {
  // E = U diag(s) V^T;  V and s^2 from the eigen-decomposition of E^T E, then E' = (E v1) v1^T / |E v1| ... with s = (1,1,0):
  // E' = u1 v1^T + u2 v2^T where (v1, v2) are the two leading right singular vectors and u_i = E v_i / s_i.
  double M[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { M[i][j] = E[0 + i] * E[0 + j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j]; V[i][j] = (i == j); }
  for (int sweep = 0; sweep < 20; ++sweep) {
    const double off = M[0][1] * M[0][1] + M[0][2] * M[0][2] + M[1][2] * M[1][2];
    if (off < 1e-30) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(M[p][q]) < 1e-300) continue;
        const double theta = (M[q][q] - M[p][p]) / (2.0 * M[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; ++k) { const double a = M[k][p], b = M[k][q]; M[k][p] = c * a - sn * b; M[k][q] = sn * a + c * b; }
        for (int k = 0; k < 3; ++k) { const double a = M[p][k], b = M[q][k]; M[p][k] = c * a - sn * b; M[q][k] = sn * a + c * b; }
        for (int k = 0; k < 3; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - sn * b; V[k][q] = sn * a + c * b; }
      }
  }
  int lo = 0;                                   // column of the smallest eigenvalue: dropped
  for (int i = 1; i < 3; ++i)
    if (M[i][i] < M[lo][lo]) lo = i;
  double out[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = 0; c < 3; ++c) {
    if (c == lo) continue;
    double u[3];
    for (int r = 0; r < 3; ++r) u[r] = E[3 * r] * V[0][c] + E[3 * r + 1] * V[1][c] + E[3 * r + 2] * V[2][c];
    const double nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (nrm < 1e-300) continue;
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) out[3 * r + k] += (u[r] / nrm) * V[k][c];
  }
  for (int k = 0; k < 9; ++k) E[k] = out[k];
}

// squared Sampson distance of (x,y) <-> (u,v) under x1^T E x0 = 0
__device__ __forceinline__ float rs_sampson2(const float (&E)[9], float x, float y, float u, float v) {
  const float a = E[0] * x + E[1] * y + E[2], b = E[3] * x + E[4] * y + E[5], c = E[6] * x + E[7] * y + E[8];
  const float d = E[0] * u + E[3] * v + E[6], f = E[1] * u + E[4] * v + E[7];
  const float e = u * a + v * b + c;
  const float den = a * a + b * b + d * d + f * f;
  return den > 1e-20f ? e * e / den : 1e30f;
}

__device__ __forceinline__ void rs_epi_row(float x, float y, float u, float v, double (&r)[9]) {
  r[0] = (double)u * x; r[1] = (double)u * y; r[2] = u; r[3] = (double)v * x; r[4] = (double)v * y; r[5] = v; r[6] = x; r[7] = y; r[8] = 1.0;
}

__global__ void __launch_bounds__(RS_THREADS) ransac_e_hyp_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                                                  const int* __restrict__ np, int n_max, float thr,
                                                                  int rounds_per_warp, uint32_t seed, RsRecord* __restrict__ best) {
  extern __shared__ float4 sPts[];                 // (x, y, u, v): already normalised image coordinates
  __shared__ RsRecord sBest[RS_WARPS];
  const int pair = blockIdx.y;
  const int n = min(np ? np[pair] : n_max, min(n_max, RS_MAX_PTS));
  RsRecord mine;
  mine.cost = 3.0e38f;
#pragma unroll
  for (int k = 0; k < 9; ++k) mine.m[k] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (n >= 8) {
    const float* q0 = p0 + (int64_t)pair * n_max * 2;
    const float* q1 = p1 + (int64_t)pair * n_max * 2;
    for (int i = threadIdx.x; i < n; i += RS_THREADS) sPts[i] = make_float4(q0[2 * i], q0[2 * i + 1], q1[2 * i], q1[2 * i + 1]);
    __syncthreads();
    const float t2 = thr * thr;
    for (int rd = 0; rd < rounds_per_warp; ++rd) {
      // every lane: one 8-point hypothesis
      const uint32_t hid = (uint32_t)(((blockIdx.x * RS_WARPS + warp) * rounds_per_warp + rd) * 32 + lane);
      uint32_t s = rs_hash(seed ^ (0x9e3779b9u * (uint32_t)(pair + 1)) ^ (0x85ebca6bu * (hid + 1)));
      int idx[8];
      for (int k = 0; k < 8; ++k) {
        bool dup;
        do {
          s = rs_hash(s + 0x6d2b79f5u);
          idx[k] = (int)(((uint64_t)s * (uint64_t)n) >> 32);
          dup = false;
          for (int j = 0; j < k; ++j) dup |= (idx[j] == idx[k]);
        } while (dup);
      }
      double M[9][9];
      for (int a = 0; a < 9; ++a)
        for (int b2 = 0; b2 < 9; ++b2) M[a][b2] = 0.0;
      for (int k = 0; k < 8; ++k) {
        const float4 c = sPts[idx[k]];
        double r[9];
        rs_epi_row(c.x, c.y, c.z, c.w, r);
        for (int a = 0; a < 9; ++a)
          for (int b2 = a; b2 < 9; ++b2) M[a][b2] += r[a] * r[b2];
      }
      for (int a = 0; a < 9; ++a)
        for (int b2 = 0; b2 < a; ++b2) M[a][b2] = M[b2][a];
      double e[9];
      rs_smallest_eigvec<9>(M, e);
      rs_project_essential(e);
      float Ef[9];
      for (int k = 0; k < 9; ++k) Ef[k] = (float)e[k];
      // the warp scores the 32 hypotheses one after the other
      for (int src = 0; src < 32; ++src) {
        float E[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) E[k] = __shfl_sync(0xffffffffu, Ef[k], src);
        float cost = 0.f;
        for (int i = lane; i < n; i += 32) {
          const float4 c = sPts[i];
          cost += fminf(rs_sampson2(E, c.x, c.y, c.z, c.w), t2);
        }
        for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
        if (cost < mine.cost) {
          mine.cost = cost;
#pragma unroll
          for (int k = 0; k < 9; ++k) mine.m[k] = E[k];
        }
      }
    }
  }
  if (lane == 0) sBest[warp] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    RsRecord b = sBest[0];
    for (int w = 1; w < RS_WARPS; ++w)
      if (sBest[w].cost < b.cost) b = sBest[w];
    best[(int64_t)pair * gridDim.x + blockIdx.x] = b;
  }
}

// Local optimisation of one essential-matrix candidate, by the whole block.  The 8-point fit of a minimal sample is noisy, so its
// consensus set is only part of the true one: RS_E_REFITS rounds of a Sampson-weighted 8-point refit (rows scaled by 1/|grad|,
// one IRLS step towards the Sampson error) over the points within mult*thr of the current model, mult shrinking 3 -> 1.  The model
// kept is the one with the lowest MSAC cost at the true threshold, so the result is never worse than the candidate.
// sE (shared, 9 floats): candidate in, optimised model out (visible to all threads on return).  Returns its MSAC cost.
__device__ float rs_e_local_opt(const float4* __restrict__ sPts, int n, float t2, float* sE) {
  __shared__ double sAcc[RS_WARPS][45];
  __shared__ float sCost[RS_WARPS];
  __shared__ float sBestE[9];
  __shared__ int sCnt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float best_cost = 3.0e38f;                         // block-uniform
  __syncthreads();
  for (int round = 0; round <= RS_E_REFITS; ++round) {
    float E[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) E[k] = sE[k];
    float cost = 0.f;                                // MSAC cost of the current model at the true threshold
    for (int i = threadIdx.x; i < n; i += RS_THREADS) {
      const float4 c = sPts[i];
      cost += fminf(rs_sampson2(E, c.x, c.y, c.z, c.w), t2);
    }
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if (lane == 0) sCost[warp] = cost;
    __syncthreads();
    float total = 0.f;
    for (int w = 0; w < RS_WARPS; ++w) total += sCost[w];      // same order in every thread
    if (total < best_cost) {
      best_cost = total;
      if (threadIdx.x < 9) sBestE[threadIdx.x] = E[threadIdx.x];
    }
    if (round == RS_E_REFITS) break;
    const float mult = round < 2 ? 3.0f : (round < 4 ? 2.0f : (round < 5 ? 1.5f : 1.0f));
    const float t2r = t2 * mult * mult;
    double acc[45];
#pragma unroll
    for (int k = 0; k < 45; ++k) acc[k] = 0.0;
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += RS_THREADS) {
      const float4 c = sPts[i];
      const float ea = E[0] * c.x + E[1] * c.y + E[2], eb = E[3] * c.x + E[4] * c.y + E[5];
      const float ed = E[0] * c.z + E[3] * c.w + E[6], ef = E[1] * c.z + E[4] * c.w + E[7];
      const float den = ea * ea + eb * eb + ed * ed + ef * ef;
      if (rs_sampson2(E, c.x, c.y, c.z, c.w) < t2r && den > 1e-20f) {
        ++cnt;
        double r[9];
        rs_epi_row(c.x, c.y, c.z, c.w, r);
        const double wgt = 1.0 / (double)den;
        int e = 0;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const double ra = r[a] * wgt;
#pragma unroll
          for (int b2 = a; b2 < 9; ++b2) acc[e++] += ra * r[b2];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 45; ++k) {
      double v = acc[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sAcc[warp][k] = v;
    }
    if (threadIdx.x == 0) sCnt = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) atomicAdd(&sCnt, cnt);
    __syncthreads();
    if (sCnt < 8) break;                             // block-uniform
    if (threadIdx.x == 0) {
      double M[9][9], vec[9];
      int e = 0;
      for (int a = 0; a < 9; ++a)
        for (int b2 = a; b2 < 9; ++b2) {
          double v = 0.0;
          for (int w = 0; w < RS_WARPS; ++w) v += sAcc[w][e];
          M[a][b2] = v; M[b2][a] = v;
          ++e;
        }
      rs_smallest_eigvec<9>(M, vec);
      rs_project_essential(vec);
      for (int k = 0; k < 9; ++k) sE[k] = (float)vec[k];
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x < 9) sE[threadIdx.x] = sBestE[threadIdx.x];
  __syncthreads();
  return best_cost;
}

// stage 2: every block's winner is optimised on its own (grid (blocks_per_pair, batch)) -- with w^8 all-inlier samples the raw
// winners are few and noisy, and which of them converges to the full consensus set is not predictable from their raw cost.
__global__ void __launch_bounds__(RS_THREADS) ransac_e_lo_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                                                 const int* __restrict__ np, int n_max, float thr,
                                                                 RsRecord* __restrict__ best) {
  extern __shared__ float4 sPts[];
  __shared__ float sE[9];
  const int pair = blockIdx.y;
  const int n = min(np ? np[pair] : n_max, min(n_max, RS_MAX_PTS));
  if (n < 8) return;
  RsRecord* rec = best + (int64_t)pair * gridDim.x + blockIdx.x;
  if (!(rec->cost < 3.0e38f)) return;                // this block found no hypothesis (block-uniform)
  const float* q0 = p0 + (int64_t)pair * n_max * 2;
  const float* q1 = p1 + (int64_t)pair * n_max * 2;
  for (int i = threadIdx.x; i < n; i += RS_THREADS) sPts[i] = make_float4(q0[2 * i], q0[2 * i + 1], q1[2 * i], q1[2 * i + 1]);
  if (threadIdx.x < 9) sE[threadIdx.x] = rec->m[threadIdx.x];
  const float cost = rs_e_local_opt(sPts, n, thr * thr, sE);
  if (threadIdx.x < 9) rec->m[threadIdx.x] = sE[threadIdx.x];
  if (threadIdx.x == 0) rec->cost = cost;
}

// stage 3: the best optimised candidate classifies the correspondences.
__global__ void __launch_bounds__(RS_THREADS) ransac_e_final_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                                                    const int* __restrict__ np, int n_max, float thr,
                                                                    const RsRecord* __restrict__ best, int blocks_per_pair,
                                                                    float* __restrict__ E_out, unsigned char* __restrict__ mask,
                                                                    int* __restrict__ n_inl) {
  __shared__ float sE[9];
  __shared__ int sCnt;
  const int pair = blockIdx.x;
  const int n_all = np ? np[pair] : n_max;
  const int n = min(n_all, min(n_max, RS_MAX_PTS));
  unsigned char* mk = mask + (int64_t)pair * n_max;
  for (int i = threadIdx.x; i < n_max; i += RS_THREADS) mk[i] = 0;
  if (n < 8) {
    if (threadIdx.x == 0) {
      n_inl[pair] = n_all < 0 ? n_all : 0;
      for (int k = 0; k < 9; ++k) E_out[pair * 9 + k] = 0.f;
    }
    return;
  }
  const float* q0 = p0 + (int64_t)pair * n_max * 2;
  const float* q1 = p1 + (int64_t)pair * n_max * 2;
  if (threadIdx.x == 0) {
    RsRecord b = best[(int64_t)pair * blocks_per_pair];
    for (int k = 1; k < blocks_per_pair; ++k) {
      const RsRecord c = best[(int64_t)pair * blocks_per_pair + k];
      if (c.cost < b.cost) b = c;
    }
    for (int k = 0; k < 9; ++k) sE[k] = b.m[k];
    sCnt = 0;
  }
  __syncthreads();
  const float t2 = thr * thr;
  const int lane = threadIdx.x & 31;
  float E[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) E[k] = sE[k];
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += RS_THREADS) {
    const bool in = rs_sampson2(E, q0[2 * i], q0[2 * i + 1], q1[2 * i], q1[2 * i + 1]) < t2;
    mk[i] = in ? 1 : 0;
    cnt += in ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) atomicAdd(&sCnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    n_inl[pair] = sCnt;
    for (int k = 0; k < 9; ++k) E_out[pair * 9 + k] = E[k];
  }
}

}  // namespace xf

extern "C" size_t xfeat_ransac_workspace_bytes(int batch, int iters) {
  const int bpp = (iters + xf::RS_WARPS * 32 - 1) / (xf::RS_WARPS * 32);   // 32 hypotheses per warp
  return xf::align_up((size_t)batch * (bpp > 0 ? bpp : 1) * sizeof(xf::RsRecord), 256);
}

extern "C" int xfeat_ransac_homography(const float* d_pts0, const float* d_pts1, const int32_t* d_n, int n_max, int batch,
                                       float thr_px, int iters, uint32_t seed, float* d_H, uint8_t* d_inliers,
                                       int32_t* d_n_inliers, void* d_ws, size_t ws_bytes, void* stream) {
  XF_REQUIRE(d_pts0 && d_pts1 && d_H && d_inliers && d_n_inliers && d_ws, "ransac_homography: null pointer");
  XF_REQUIRE(batch > 0 && batch <= 65535 && n_max > 0 && n_max <= xf::RS_MAX_PTS && iters > 0 && thr_px > 0.f,
             "ransac_homography: bad arguments (n_max <= %d)", xf::RS_MAX_PTS);
  XF_REQUIRE(ws_bytes >= xfeat_ransac_workspace_bytes(batch, iters), "ransac_homography: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int bpp = xf::cdiv(iters, xf::RS_WARPS * 32);
  const int hyp_per_warp = xf::cdiv(iters, bpp * xf::RS_WARPS);
  const size_t smem = (size_t)n_max * sizeof(float4);
  XF_DYN_SMEM(xf::ransac_h_hyp_kernel, smem);
  XF_DYN_SMEM(xf::ransac_h_final_kernel, smem);
  xf::ransac_h_hyp_kernel<<<dim3(bpp, batch), xf::RS_THREADS, smem, st>>>(d_pts0, d_pts1, d_n, n_max, thr_px, hyp_per_warp, seed,
                                                                          (xf::RsRecord*)d_ws);
  XF_LAUNCH_CHECK();
  xf::ransac_h_final_kernel<<<batch, xf::RS_THREADS, smem, st>>>(d_pts0, d_pts1, d_n, n_max, thr_px, (const xf::RsRecord*)d_ws, bpp,
                                                                 d_H, d_inliers, d_n_inliers);
  XF_LAUNCH_CHECK();
  return XF_OK;
}

extern "C" int xfeat_ransac_essential(const float* d_x0, const float* d_x1, const int32_t* d_n, int n_max, int batch, float thr,
                                      int iters, uint32_t seed, float* d_E, uint8_t* d_inliers, int32_t* d_n_inliers, void* d_ws,
                                      size_t ws_bytes, void* stream) {
  XF_REQUIRE(d_x0 && d_x1 && d_E && d_inliers && d_n_inliers && d_ws, "ransac_essential: null pointer");
  XF_REQUIRE(batch > 0 && batch <= 65535 && n_max > 0 && n_max <= xf::RS_MAX_PTS && iters > 0 && thr > 0.f,
             "ransac_essential: bad arguments (n_max <= %d)", xf::RS_MAX_PTS);
  XF_REQUIRE(ws_bytes >= xfeat_ransac_workspace_bytes(batch, iters), "ransac_essential: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int bpp = xf::cdiv(iters, xf::RS_WARPS * 32);          // one round = 32 hypotheses per warp
  const int rounds = xf::cdiv(iters, bpp * xf::RS_WARPS * 32);
  const size_t smem = (size_t)n_max * sizeof(float4);
  XF_DYN_SMEM(xf::ransac_e_hyp_kernel, smem);
  XF_DYN_SMEM(xf::ransac_e_lo_kernel, smem);
  xf::ransac_e_hyp_kernel<<<dim3(bpp, batch), xf::RS_THREADS, smem, st>>>(d_x0, d_x1, d_n, n_max, thr, rounds, seed,
                                                                          (xf::RsRecord*)d_ws);
  XF_LAUNCH_CHECK();
  xf::ransac_e_lo_kernel<<<dim3(bpp, batch), xf::RS_THREADS, smem, st>>>(d_x0, d_x1, d_n, n_max, thr, (xf::RsRecord*)d_ws);
  XF_LAUNCH_CHECK();
  xf::ransac_e_final_kernel<<<batch, xf::RS_THREADS, 0, st>>>(d_x0, d_x1, d_n, n_max, thr, (const xf::RsRecord*)d_ws, bpp, d_E,
                                                                 d_inliers, d_n_inliers);
  XF_LAUNCH_CHECK();
  return XF_OK;
}
