// Dense tail of a mini-batch GGAD training batch, fused (src/graphsage.py:421-450 GCNEncoder, :174,:197-198,:234-258 GCN):
//
//   given  C   [B,h]  = ReLU(W . to_feats^T)^T        (projection, ggad_dense_matmul)
//          ego [B,h]  = mask_row . emb_U               (ego-neighbor mean, ggad_gather_reduce)
//          fc  [h,h], weight [h], labels [B] in {0,1}
//   afn[b]   = ReLU(fc . ego[b])                                        outlier generation            (:430)
//   R[p]     = label[src[p]] == 1 ? afn[src[p]] : C[src[p]]             combined_all^T, src = stable sort by label (:450)
//   s[p]     = weight . R[p] ;  bce[p] = BCEWithLogits(s[p], label[p])  labels stay in batch order     (:174,:246)
//   cos[p]   = <R[p], ego[p]> / (max(|R[p]|,eps) max(|ego[p]|,eps))     column p against ego row p     (:234)
//   margin   = max(0, 1 - (mean_{label 0} cos - mean_{label 1} cos))                                   (:236-240)
//   rec      = mean_{label[b] = 1} |C[b] - afn[b]|_2                                                   (:197-198)
//   total    = mean bce + margin + 0.1 rec                                                             (:258)
//
// One CTA per batch position (h threads), a one-CTA reduction for the four scalars, and the hand-derived backward in the
// same shape: every gradient row has one writer, or exactly two commutative atomic contributions -> deterministic.
// Replaces ~40 (forward) + ~60 (backward) elementwise / reduction launches of the autograd formulation.
#include <math.h>

#include "common.cuh"

namespace ggad {

constexpr float kCosEps = 1e-8f;  // torch.cosine_similarity's eps

struct TailArgs {
  const float *C, *ego, *fc, *w;
  const int64_t* lab;
  int32_t B, h;
  int64_t ldc, lde;
  // forward outputs / backward inputs
  float *R, *apre_src, *apre_own, *s, *bce, *cos, *dist, *nrm;  // nrm[2p] = |R[p]|, nrm[2p+1] = |ego[p]|
  int32_t* src;
  float* out;  // [8]: total, cls, margin, rec, margin active (0/1), n0, n1, -
  // backward
  const float* g_total;
  float *dC, *dapre, *dego, *ds, *dw;
  int64_t lddc;
};

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += s_red[i];  // fixed order
  return t;
}

// position p -> source row of combined_all: label-0 rows in batch order first, then label-1 rows (stable)
__device__ __forceinline__ int source_row(const int64_t* __restrict__ lab, int B, int p, int* s_tmp) {
  if (threadIdx.x == 0) {
    int n0 = 0;
    for (int b = 0; b < B; ++b) n0 += (lab[b] == 0);
    const int64_t cls = (p < n0) ? 0 : 1;
    int want = (p < n0) ? p : p - n0, q = -1;
    for (int b = 0; b < B; ++b)
      if (lab[b] == cls) {
        if (want == 0) { q = b; break; }
        --want;
      }
    if (q < 0) q = 0;   // labels outside {0,1}: the host side refuses them; never index out of bounds
    *s_tmp = q;
  }
  __syncthreads();
  return *s_tmp;
}

extern __shared__ float tail_smem[];

__global__ void tail_rows_fwd(const __grid_constant__ TailArgs a) {
  const int p = blockIdx.x, t = threadIdx.x, h = a.h;
  float* s_ego_q = tail_smem;          // [h]
  float* s_ego_p = tail_smem + h;      // [h]
  float* s_red = tail_smem + 2 * h;    // [32]
  __shared__ int s_q;
  const int q = source_row(a.lab, a.B, p, &s_q);
  if (t == 0) a.src[p] = q;
  const bool live = t < h;
  const bool ab_q = a.lab[q] == 1, ab_p = a.lab[p] == 1;
  if (live) {
    s_ego_q[t] = a.ego[int64_t(q) * a.lde + t];
    s_ego_p[t] = a.ego[int64_t(p) * a.lde + t];
  }
  __syncthreads();
  float r = 0.f;
  if (live) {
    if (ab_q) {
      float acc = 0.f;
      const float* frow = a.fc + int64_t(t) * h;
      for (int j = 0; j < h; ++j) acc = fmaf(__ldg(frow + j), s_ego_q[j], acc);
      a.apre_src[int64_t(p) * h + t] = acc;
      r = fmaxf(acc, 0.f);
    } else {
      r = a.C[int64_t(q) * a.ldc + t];
    }
    a.R[int64_t(p) * h + t] = r;
  }
  const float e = live ? s_ego_p[t] : 0.f;
  const float sc = block_sum(live ? __ldg(a.w + t) * r : 0.f, s_red);
  const float rr = block_sum(r * r, s_red);
  const float ee = block_sum(e * e, s_red);
  const float re = block_sum(r * e, s_red);
  // reconstruction term of row p itself (only rows with label 1): |C[p] - ReLU(fc ego[p])|
  float d2 = 0.f;
  if (ab_p && live) {
    float acc = 0.f;
    const float* frow = a.fc + int64_t(t) * h;
    for (int j = 0; j < h; ++j) acc = fmaf(__ldg(frow + j), s_ego_p[j], acc);
    a.apre_own[int64_t(p) * h + t] = acc;
    const float df = a.C[int64_t(p) * a.ldc + t] - fmaxf(acc, 0.f);
    d2 = df * df;
  }
  d2 = block_sum(d2, s_red);
  if (t == 0) {
    const float y = float(a.lab[p] == 1);
    a.s[p] = sc;
    a.bce[p] = fmaxf(sc, 0.f) - sc * y + log1pf(expf(-fabsf(sc)));
    const float nr = sqrtf(rr), ne = sqrtf(ee);
    a.nrm[2 * p] = nr;
    a.nrm[2 * p + 1] = ne;
    a.cos[p] = re / (fmaxf(nr, kCosEps) * fmaxf(ne, kCosEps));
    a.dist[p] = ab_p ? sqrtf(d2) : 0.f;
  }
}

__global__ void tail_reduce_fwd(const __grid_constant__ TailArgs a) {
  __shared__ float s_red[32];
  float bce = 0.f, c0 = 0.f, c1 = 0.f, n0 = 0.f, n1 = 0.f, ds = 0.f;
  for (int p = threadIdx.x; p < a.B; p += blockDim.x) {
    const int64_t l = a.lab[p];
    bce += a.bce[p];
    if (l == 0) { c0 += a.cos[p]; n0 += 1.f; }
    if (l == 1) { c1 += a.cos[p]; n1 += 1.f; ds += a.dist[p]; }
  }
  bce = block_sum(bce, s_red); c0 = block_sum(c0, s_red); c1 = block_sum(c1, s_red);
  n0 = block_sum(n0, s_red); n1 = block_sum(n1, s_red); ds = block_sum(ds, s_red);
  if (threadIdx.x == 0) {
    const float cls = bce / float(a.B);
    const float m = 1.f - (c0 / n0 - c1 / n1);      // 0/0 -> NaN like the reference's mean of an empty selection
    const float margin = fmaxf(m, 0.f) + (m != m ? m : 0.f);
    const float rec = ds / n1;
    a.out[0] = cls + margin + 0.1f * rec;
    a.out[1] = cls; a.out[2] = margin; a.out[3] = rec;
    a.out[4] = (m > 0.f) ? 1.f : 0.f; a.out[5] = n0; a.out[6] = n1; a.out[7] = 0.f;
  }
}

__global__ void tail_rows_bwd(const __grid_constant__ TailArgs a) {
  const int p = blockIdx.x, t = threadIdx.x, h = a.h;
  float* s_vec = tail_smem;           // [h] gradient w.r.t. a pre-activation row (matvec operand)
  float* s_red = tail_smem + h;       // [32]
  const bool live = t < h;
  const int q = a.src[p];
  const bool ab_q = a.lab[q] == 1, ab_p = a.lab[p] == 1;
  const float g = __ldg(a.g_total);
  const float n0 = a.out[5], n1 = a.out[6], active = a.out[4];
  const float sc = a.s[p], y = float(ab_p);
  const float ds = g * (1.f / (1.f + expf(-sc)) - y) / float(a.B);
  if (t == 0) a.ds[p] = ds;
  const float gcos = g * active * (a.lab[p] == 0 ? -1.f / n0 : (ab_p ? 1.f / n1 : 0.f));
  const float nr = a.nrm[2 * p], ne = a.nrm[2 * p + 1], cs = a.cos[p];
  const float cnr = fmaxf(nr, kCosEps), cne = fmaxf(ne, kCosEps);
  const float r = live ? a.R[int64_t(p) * h + t] : 0.f;
  const float e = live ? a.ego[int64_t(p) * a.lde + t] : 0.f;
  // d cos / d R and d cos / d ego (a clamped norm is a constant, like autograd through clamp_min)
  const float dcr = e / (cnr * cne) - (nr > kCosEps ? cs * r / (nr * nr) : 0.f);
  const float dce = r / (cnr * cne) - (ne > kCosEps ? cs * e / (ne * ne) : 0.f);
  const float dR = live ? ds * __ldg(a.w + t) + gcos * dcr : 0.f;
  float dego_p = gcos * dce;          // cosine contribution to d ego[p]
  // ---- scatter dR to its source row q ----
  float da_src = 0.f;
  if (live) {
    if (ab_q) da_src = (a.apre_src[int64_t(p) * h + t] > 0.f) ? dR : 0.f;
    else a.dC[int64_t(q) * a.lddc + t] = dR;                      // single writer: position pos[q] = p
  }
  if (ab_q) {                                                     // d ego[q] += fc^T da_src ; d apre[q] += da_src
    __syncthreads();
    if (live) s_vec[t] = da_src;
    __syncthreads();
    if (live) {
      float acc = 0.f;
      for (int i = 0; i < h; ++i) acc = fmaf(__ldg(a.fc + int64_t(i) * h + t), s_vec[i], acc);
      atomicAdd(a.dego + int64_t(q) * h + t, acc);
      atomicAdd(a.dapre + int64_t(q) * h + t, da_src);
    }
  }
  // ---- reconstruction term of row p itself ----
  if (ab_p) {
    const float grec = g * 0.1f / n1;
    const float dist = a.dist[p];
    float da_own = 0.f;
    if (live) {
      const float ap = a.apre_own[int64_t(p) * h + t];
      const float df = a.C[int64_t(p) * a.ldc + t] - fmaxf(ap, 0.f);
      const float gd = dist > 0.f ? grec * df / dist : 0.f;
      a.dC[int64_t(p) * a.lddc + t] = gd;                         // single writer: a label-1 row is nobody's C source
      da_own = (ap > 0.f) ? -gd : 0.f;
    }
    __syncthreads();
    if (live) s_vec[t] = da_own;
    __syncthreads();
    if (live) {
      float acc = 0.f;
      for (int i = 0; i < h; ++i) acc = fmaf(__ldg(a.fc + int64_t(i) * h + t), s_vec[i], acc);
      dego_p += acc;
      atomicAdd(a.dapre + int64_t(p) * h + t, da_own);
    }
  }
  if (live) atomicAdd(a.dego + int64_t(p) * h + t, dego_p);
  (void)s_red;
}

// dw[t] = sum_p ds[p] R[p,t]  (fixed order)
__global__ void tail_dw_kernel(const __grid_constant__ TailArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.h) return;
  float acc = 0.f;
  for (int p = 0; p < a.B; ++p) acc = fmaf(a.ds[p], a.R[int64_t(p) * a.h + t], acc);
  a.dw[t] = acc;
}

static int fill(TailArgs& a, const ggad_tail_desc_t* d) {
  GGAD_REQUIRE(d != nullptr, GGAD_ERR_INVALID, "minibatch_tail: null descriptor");
  GGAD_REQUIRE(d->batch > 0 && d->h > 0 && d->h <= 256, GGAD_ERR_UNSUPPORTED, "minibatch_tail: batch must be > 0 and 0 < h <= 256");
  GGAD_REQUIRE(d->combined && d->ego && d->fc && d->weight && d->labels && d->rows && d->apre_src && d->apre_own && d->scores &&
                   d->bce && d->cos && d->dist && d->norms && d->src && d->out,
               GGAD_ERR_INVALID, "minibatch_tail: null pointer");
  GGAD_REQUIRE(d->ld_combined >= d->h && d->ld_ego >= d->h, GGAD_ERR_INVALID, "minibatch_tail: leading dimension < h");
  a.C = d->combined; a.ego = d->ego; a.fc = d->fc; a.w = d->weight; a.lab = d->labels; a.B = d->batch; a.h = d->h;
  a.ldc = d->ld_combined; a.lde = d->ld_ego;
  a.R = d->rows; a.apre_src = d->apre_src; a.apre_own = d->apre_own; a.s = d->scores; a.bce = d->bce; a.cos = d->cos;
  a.dist = d->dist; a.nrm = d->norms; a.src = d->src; a.out = d->out;
  a.g_total = d->grad_total; a.dC = d->d_combined; a.dapre = d->d_apre; a.dego = d->d_ego; a.ds = d->d_scores; a.dw = d->d_weight;
  a.lddc = d->ld_d_combined;
  return GGAD_OK;
}

int minibatch_tail_fwd_impl(const ggad_tail_desc_t* d, cudaStream_t st) {
  TailArgs a;
  int rc = fill(a, d);
  if (rc != GGAD_OK) return rc;
  const int threads = ((a.h + 31) / 32) * 32;
  tail_rows_fwd<<<a.B, threads, (2 * a.h + 32) * sizeof(float), st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  tail_reduce_fwd<<<1, 256, 0, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return GGAD_OK;
}

int minibatch_tail_bwd_impl(const ggad_tail_desc_t* d, cudaStream_t st) {
  TailArgs a;
  int rc = fill(a, d);
  if (rc != GGAD_OK) return rc;
  GGAD_REQUIRE(d->grad_total && d->d_combined && d->d_apre && d->d_ego && d->d_scores && d->d_weight && d->ld_d_combined >= d->h,
               GGAD_ERR_INVALID, "minibatch_tail_bwd: null gradient pointer");
  // d_combined rows all have a writer; d_apre / d_ego are accumulated with (<= 2, commutative) atomics
  GGAD_CUDA_OK(cudaMemsetAsync(a.dapre, 0, size_t(a.B) * a.h * 4, st));
  GGAD_CUDA_OK(cudaMemsetAsync(a.dego, 0, size_t(a.B) * a.h * 4, st));
  const int threads = ((a.h + 31) / 32) * 32;
  tail_rows_bwd<<<a.B, threads, (a.h + 32) * sizeof(float), st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  tail_dw_kernel<<<(a.h + 127) / 128, 128, 0, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return GGAD_OK;
}

}  // namespace ggad
