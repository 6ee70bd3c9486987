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

// best_vec[p,:] = feats[n*,p,:] where n* = clip of this call that holds best_key[p] (keys carry the global clip index);
// prototypes whose best clip lies outside [offset, offset+N) keep their row.  One block per prototype.
__global__ void push_capture_dense_kernel(const unsigned long long* __restrict__ best_key, int P, int D, long long offset,
                                          int N, const float* __restrict__ feats, float* __restrict__ best_vec) {
  const int p = blockIdx.x;
  const unsigned long long key = best_key[p] ^ PASN_KEY_SIGN;
  if (key == PASN_KEY_NONE) return;
  const long long idx = (long long)(key & 0xFFFFFFFFull);
  if (idx < offset || idx >= offset + N) return;
  const float* src = feats + ((size_t)(idx - offset) * P + p) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) best_vec[(size_t)p * D + d] = src[d];
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
  if (push && push->best_vec) {   // winner capture in the same pass (push_abs_revision.py:299-302)
    push_capture_dense_kernel<<<P, 128, 0, st>>>(reinterpret_cast<const unsigned long long*>(push->best_key), P, D,
                                                 (long long)push->global_offset, N, feats, push->best_vec);
    PASN_LAUNCH_CHECK();
    count_launch();
  }
  return PASN_OK;
}

// ---------------------------------------------------------------------------------------------
// push bookkeeping
// ---------------------------------------------------------------------------------------------
__global__ void push_init_kernel(unsigned long long* k, int P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) k[i] = PASN_KEY_NONE ^ PASN_KEY_SIGN;   // INT64_MAX in the signed-order global format
}
// decode: index / distance of the winner per prototype
__global__ void push_decode_kernel(const unsigned long long* k, int P, int64_t* index, float* distance) {
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
}
// merge of R push records [keys P x u64 | vecs P x D x f32]: per prototype (one block) the smallest signed key wins
__global__ void push_reduce_kernel(const unsigned char* __restrict__ gathered, size_t rec_bytes, int R, int P, int D,
                                   int64_t* __restrict__ index, float* __restrict__ distance, int32_t* __restrict__ valid,
                                   float* __restrict__ vec) {
  const int p = blockIdx.x;
  long long best = 0x7FFFFFFFFFFFFFFFll;
  int rb = 0;
  for (int r = 0; r < R; ++r) {
    const long long k = reinterpret_cast<const long long*>(gathered + (size_t)r * rec_bytes)[p];
    if (k < best) { best = k; rb = r; }
  }
  const unsigned long long key = (unsigned long long)best ^ PASN_KEY_SIGN;
  const bool ok = key != PASN_KEY_NONE;
  if (threadIdx.x == 0) {
    index[p] = ok ? (long long)(key & 0xFFFFFFFFull) : -1;
    if (distance) distance[p] = ok ? f32_from_orderable((uint32_t)(key >> 32)) : __int_as_float(0x7f800000);
    if (valid) valid[p] = ok ? 1 : 0;
  }
  if (ok && vec) {
    const float* src = reinterpret_cast<const float*>(gathered + (size_t)rb * rec_bytes + (size_t)P * 8) + (size_t)p * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) vec[(size_t)p * D + d] = src[d];
  }
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
extern "C" size_t pasn_push_record_bytes(int32_t P, int32_t D) {
  return (P > 0 && D > 0) ? (size_t)P * 8 + (size_t)P * D * 4 : 0;
}
extern "C" int pasn_push_decode(const uint64_t* best_key, int32_t P, int64_t* index, float* distance, void* stream) {
  if (!best_key || !index || P <= 0) return PASN_ERR_INVALID;
  push_decode_kernel<<<ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(best_key), P, index, distance);
  PASN_LAUNCH_CHECK();
  count_launch();
  return PASN_OK;
}
extern "C" int pasn_push_reduce(const void* gathered, int32_t R, int32_t P, int32_t D, int64_t* index, float* distance,
                                int32_t* valid, float* vec, void* stream) {
  if (!gathered || !index || R <= 0 || P <= 0 || D <= 0) return PASN_ERR_INVALID;
  if (((uintptr_t)gathered & 7) != 0) return PASN_ERR_ALIGN;
  push_reduce_kernel<<<P, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned char*>(gathered),
                                                          pasn_push_record_bytes(P, D), R, P, D, index, distance, valid, vec);
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
