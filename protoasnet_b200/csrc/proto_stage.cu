// Prototype stage: pooled features -> cosine similarity vs prototype vectors -> (cos+1)/2 -> last-layer logits
// -> push distance 1-s -> class-restricted running argmin keys.
//
// Reference arithmetic being restated (fp32 throughout):
//   nn.CosineSimilarity(dim=2, eps=1e-8): each norm clamped separately, normalise, then dot
//                                                       src/models/Video_XProtoNet.py:65, :90-92
//   similarity = (similarity + 1) / 2.0                 :93
//   logits = last_layer(similarity)  (no bias)          :96
//   distance = 1 - similarity                           :130
//   running argmin with class mask                      src/utils/push_abs_revision.py:288-307
// Bit-exact push indices need this exact rounding chain: argmin is taken over d = 1-(cos+1)/2, never over
// raw cos.  Ties resolve to the lowest global index through the packed key.
#include "common.cuh"

namespace pasn {

constexpr int PS_THREADS = 256;
constexpr int PS_WARPS = PS_THREADS / 32;

// One block walks clips n = blockIdx.x, blockIdx.x + gridDim.x, ...; warp w handles prototypes w, w+8, ...
// dynamic smem: float s_sim[P]; unsigned long long s_key[P] (push only)
__global__ void __launch_bounds__(PS_THREADS) proto_stage_kernel(
    const float* __restrict__ feats, const float* __restrict__ protos, const float* __restrict__ last_layer, int N,
    int P, int D, int K, float* __restrict__ logits, float* __restrict__ sim, float* __restrict__ dist,
    const int64_t* __restrict__ labels, const int32_t* __restrict__ proto_class, long long global_offset,
    unsigned long long* __restrict__ best_key) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem_raw);
  float* s_sim = reinterpret_cast<float*>(smem_raw + (size_t)P * sizeof(unsigned long long));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool do_push = best_key != nullptr;
  if (do_push)
    for (int p = threadIdx.x; p < P; p += PS_THREADS) s_key[p] = PASN_KEY_NONE;
  __syncthreads();

  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const long long label = do_push ? labels[n] : 0;
    for (int p = warp; p < P; p += PS_WARPS) {
      const float* f = feats + ((size_t)n * P + p) * D;
      const float* v = protos + (size_t)p * D;
      float ff = 0.f, vv = 0.f;
      for (int d = lane; d < D; d += 32) {
        float a = f[d], b = v[d];
        ff = fmaf(a, a, ff);
        vv = fmaf(b, b, vv);
      }
      ff = warp_sum(ff);
      vv = warp_sum(vv);
      const float nf = fmaxf(sqrtf(ff), 1e-8f), nv = fmaxf(sqrtf(vv), 1e-8f);
      float dot = 0.f;
      for (int d = lane; d < D; d += 32) dot = fmaf(f[d] / nf, v[d] / nv, dot);
      dot = warp_sum(dot);
      if (lane == 0) {
        const float s = (dot + 1.0f) / 2.0f;
        const float dd = 1.0f - s;
        s_sim[p] = s;
        sim[(size_t)n * P + p] = s;
        if (dist) dist[(size_t)n * P + p] = dd;
        if (do_push) {
          const int pc = proto_class[p];
          if (pc < 0 || (long long)pc == label) {
            unsigned long long key = pack_key(dd, (uint32_t)(global_offset + n));
            if (key < s_key[p]) s_key[p] = key;  // prototype p is owned by exactly one warp of this block
          }
        }
      }
    }
    __syncthreads();
    for (int k = warp; k < K; k += PS_WARPS) {
      float acc = 0.f;
      for (int p = lane; p < P; p += 32) acc = fmaf(s_sim[p], last_layer[(size_t)k * P + p], acc);
      acc = warp_sum(acc);
      if (lane == 0) logits[(size_t)n * K + k] = acc;
    }
    __syncthreads();
  }
  if (do_push)
    for (int p = threadIdx.x; p < P; p += PS_THREADS)
      if (s_key[p] != PASN_KEY_NONE) key_atomic_min_global(&best_key[p], s_key[p]);
}

int launch_proto_stage(const float* feats, const float* protos, const float* last_layer, int N, int P, int D, int K,
                       float* logits, float* sim, float* dist, const pasn_push_args* push, cudaStream_t st) {
  if (N <= 0) return PASN_OK;
  size_t smem = (size_t)P * (sizeof(unsigned long long) + sizeof(float));
  if (smem > 200 * 1024) return PASN_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    if (cudaFuncSetAttribute(proto_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return PASN_ERR_CUDA;
  int grid = N < 148 * 8 ? N : 148 * 8;
  proto_stage_kernel<<<grid, PS_THREADS, smem, st>>>(
      feats, protos, last_layer, N, P, D, K, logits, sim, dist, push ? push->labels : nullptr,
      push ? push->proto_class : nullptr, push ? (long long)push->global_offset : 0,
      push ? reinterpret_cast<unsigned long long*>(push->best_key) : nullptr);
  PASN_LAUNCH_CHECK();
  count_launch();
  return PASN_OK;
}

// ---------------------------------------------------------------------------------------------
// push bookkeeping
// ---------------------------------------------------------------------------------------------
__global__ void push_init_kernel(unsigned long long* k, int P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) k[i] = PASN_KEY_NONE ^ PASN_KEY_SIGN;   // INT64_MAX in the signed-order global format
}
// decode + ownership in one launch: index / distance of the winner, whether this rank's range [lo, hi) owns it, and the
// (clamped) local index to re-fetch it from
__global__ void push_select_kernel(const unsigned long long* k, int P, long long lo, long long hi, int64_t* index,
                                   float* distance, int64_t* local_index, int32_t* own, int32_t* valid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const unsigned long long key = k[i] ^ PASN_KEY_SIGN;
  long long idx = -1;
  float d = __int_as_float(0x7f800000);
  if (key != PASN_KEY_NONE) {
    idx = (long long)(key & 0xFFFFFFFFull);
    d = f32_from_orderable((uint32_t)(key >> 32));
  }
  index[i] = idx;
  if (distance) distance[i] = d;
  if (valid) valid[i] = idx >= 0;
  if (own) own[i] = idx >= lo && idx < hi;
  if (local_index) {
    long long l = idx < lo ? lo : (idx >= hi ? hi - 1 : idx);
    local_index[i] = hi > lo ? l - lo : 0;
  }
}
// vec[p,:] = own[p] ? feats[p,p,:] : 0   (feats = push_forward output over the P re-fetched winner clips)
__global__ void push_collect_kernel(const float* feats, const int32_t* own, float* vec, int P, int D) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * D) return;
  const int p = (int)(i / D), d = (int)(i - (long long)p * D);
  vec[i] = own[p] ? feats[((size_t)p * P + p) * D + d] : 0.f;
}
__global__ void push_write_kernel(float* protos, const float* vec, const int32_t* valid, int P, int D) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * D) return;
  if (valid[i / D]) protos[i] = vec[i];
}

}  // namespace pasn

using namespace pasn;

extern "C" int pasn_push_init(uint64_t* best_key, int32_t P, void* stream) {
  if (!best_key || P <= 0) return PASN_ERR_INVALID;
  push_init_kernel<<<ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(best_key), P);
  PASN_LAUNCH_CHECK();
  count_launch();
  return PASN_OK;
}
extern "C" int pasn_push_decode(const uint64_t* best_key, int32_t P, int64_t* index, float* distance, void* stream) {
  if (!best_key || !index || P <= 0) return PASN_ERR_INVALID;
  push_select_kernel<<<ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(best_key), P, 0, 0, index, distance, nullptr, nullptr, nullptr);
  PASN_LAUNCH_CHECK();
  count_launch();
  return PASN_OK;
}
extern "C" int pasn_push_select(const uint64_t* best_key, int32_t P, int64_t lo, int64_t hi, int64_t* index,
                                float* distance, int64_t* local_index, int32_t* own, int32_t* valid, void* stream) {
  if (!best_key || !index || P <= 0 || hi < lo) return PASN_ERR_INVALID;
  push_select_kernel<<<ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(best_key), P, lo, hi, index, distance, local_index, own, valid);
  PASN_LAUNCH_CHECK();
  count_launch();
  return PASN_OK;
}
extern "C" int pasn_push_collect(const float* feats, const int32_t* own, float* vec, int32_t P, int32_t D, void* stream) {
  if (!feats || !own || !vec || P <= 0 || D <= 0) return PASN_ERR_INVALID;
  long long n = (long long)P * D;
  push_collect_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(feats, own, vec, P, D);
  PASN_LAUNCH_CHECK();
  count_launch();
  return PASN_OK;
}
extern "C" int pasn_push_write_prototypes(float* prototypes, const float* vec, const int32_t* valid, int32_t P,
                                          int32_t D, void* stream) {
  if (!prototypes || !vec || !valid || P <= 0 || D <= 0) return PASN_ERR_INVALID;
  long long n = (long long)P * D;
  push_write_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prototypes, vec, valid, P, D);
  PASN_LAUNCH_CHECK();
  count_launch();
  return PASN_OK;
}
