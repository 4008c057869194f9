// sketch.cu — MashMap 3.5 query-fragment sketch on the GPU (sm_100a), batched over fragments.
//
// Replaces skch::CommonFunc::sketchSequence (src/map/include/commonFunc.hpp:217-323) as called per
// fragment by MappingCore::getSeedHits (src/map/include/mappingCore.hpp:61-76):
//   * makeUpperCaseAndValidDNA (commonFunc.hpp:110-142) fused into the fragment load;
//   * canonical k-mer hash = min(MurmurHash3_x64_128(fwd k-mer ASCII, seed 42).lo64,
//     MurmurHash3_x64_128(revcomp ASCII).lo64) (commonFunc.hpp:173-182,260-275; murmur3.h:226-303),
//     palindromes (hashFwd == hashBwd) and k-mers covering a non-ACGT base are dropped;
//   * bottom-s DISTINCT hashes with first / last position and the sign of the +-1 strand tally,
//     ascending by hash (closed form of the heap + hash-map loop at :277-321).
//
// B200 mapping: one CTA per fragment; the fragment is staged once in shared memory, every thread
// hashes k-mers from shared memory (two 64-bit Murmur3 evaluations per position, all in registers),
// the (hash, position) pairs are bitonic-sorted in shared memory, run heads are ranked with a
// ballot/popc block scan and the first s runs are folded and written as 32-byte MinmerInfo records.
#include "wfb_rt.h"
#include "../../include/wfmash_b200.h"

#include <string.h>
#include <string>
#include <vector>

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */
void wfb_count_launch_();
struct SkErr { SkErr& operator=(const std::string& s) { wfb_set_last_error_(s); return *this; } SkErr& operator=(const char* s) { wfb_set_last_error_(s); return *this; } };
static SkErr sk_last_error;

#ifndef WFB_EMU
#define SK_CHECK(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      sk_last_error = std::string(#call) + ": " + cudaGetErrorString(e_);                \
      rc = (e_ == cudaErrorMemoryAllocation) ? WFB_ENOMEM : WFB_ECUDA;                   \
      goto done;                                                                         \
    }                                                                                    \
  } while (0)
#endif



#include "sketch_kernels.h"
#include "wfb_pool.h" /* this file's cudaMalloc / cudaFree go through the library's device-memory pool */

/* dynamic shared memory: keys[N] u64 | vals[N] u32 | seq[len_pad] u8 */
WFB_KERNEL(wfb_sketch_kernel, const uint8_t* seq_base, const wfb_frag_t* frags, int nfrags, int ksize, int ssize, int npow2_max,
           wfb_minmer_t* out, int32_t* out_count
#ifdef WFB_EMU
           , unsigned char* smem_emu
#endif
) {
  WFB_KERNEL_PROLOGUE
#ifndef WFB_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
#else
  unsigned char* smem = smem_emu;
#endif
  WFB_SHARED int sh_warp[32];
  for (int f = bid; f < nfrags; f += nblocks) {
    const wfb_frag_t fr = frags[f];
    WFB_SYNC();
    const int cnt = sk_sketch_block(smem, sh_warp, seq_base, fr, ksize, ssize, npow2_max, out + (size_t)f * ssize);
    if (WFB_TID == 0) out_count[f] = cnt;
  }
}

extern "C" int wfb_sketch_fragments(int device, const char* seq_base, int64_t seq_bytes, const wfb_frag_t* frags, int32_t n,
                                    int32_t kmer_size, int32_t sketch_size, wfb_minmer_t* out, int32_t* out_count,
                                    double* kernel_ms) {
  if (n < 0 || kmer_size <= 0 || kmer_size > 32 || sketch_size <= 0 || (n > 0 && (!seq_base || !frags || !out || !out_count))) {
    sk_last_error = "bad argument";
    return WFB_EINVAL;
  }
  if (kernel_ms) *kernel_ms = 0.0;
  if (n == 0) return WFB_OK;
  int maxlen = 0;
  for (int i = 0; i < n; ++i) {
    if (frags[i].len < 0 || frags[i].seq_offset < 0 || frags[i].seq_offset + frags[i].len > seq_bytes) {
      sk_last_error = "fragment out of range";
      return WFB_EINVAL;
    }
    if (frags[i].len > maxlen) maxlen = frags[i].len;
  }
  int npow2 = 1;
  while (npow2 < maxlen - kmer_size + 1) npow2 <<= 1;
  const size_t smem = (size_t)npow2 * 12 + (((size_t)maxlen + 15) & ~(size_t)15) + 16;
  if (smem > 200 * 1024) { sk_last_error = "fragment too long for the shared-memory sketch kernel (max ~16 kb)"; return WFB_EINVAL; }
  int rc = WFB_OK;
#ifndef WFB_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    sk_last_error = "no CUDA device (this library has no CPU path)";
    return WFB_ENODEV;
  }
  uint8_t* d_seq = nullptr;
  wfb_frag_t* d_frags = nullptr;
  wfb_minmer_t* d_out = nullptr;
  int32_t* d_cnt = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  SK_CHECK(cudaSetDevice(device));
  SK_CHECK(cudaMalloc(&d_seq, (size_t)seq_bytes + 16));
  SK_CHECK(cudaMalloc(&d_frags, sizeof(wfb_frag_t) * (size_t)n));
  SK_CHECK(cudaMalloc(&d_out, sizeof(wfb_minmer_t) * (size_t)n * sketch_size));
  SK_CHECK(cudaMalloc(&d_cnt, sizeof(int32_t) * (size_t)n));
  SK_CHECK(cudaMemcpy(d_seq, seq_base, (size_t)seq_bytes, cudaMemcpyHostToDevice));
  SK_CHECK(cudaMemcpy(d_frags, frags, sizeof(wfb_frag_t) * (size_t)n, cudaMemcpyHostToDevice));
  SK_CHECK(cudaMemset(d_out, 0, sizeof(wfb_minmer_t) * (size_t)n * sketch_size));
  SK_CHECK(cudaFuncSetAttribute(wfb_sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SK_CHECK(cudaEventCreate(&e0));
  SK_CHECK(cudaEventCreate(&e1));
  {
    cudaDeviceProp prop;
    SK_CHECK(cudaGetDeviceProperties(&prop, device));
    const int grid = n < prop.multiProcessorCount * 8 ? n : prop.multiProcessorCount * 8;
    SK_CHECK(cudaEventRecord(e0, 0));
    wfb_sketch_kernel<<<grid, 256, smem, 0>>>(d_seq, d_frags, n, kmer_size, sketch_size, npow2, d_out, d_cnt);
    wfb_count_launch_();
    SK_CHECK(cudaEventRecord(e1, 0));
    SK_CHECK(cudaGetLastError());
    SK_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
  }
  SK_CHECK(cudaMemcpy(out, d_out, sizeof(wfb_minmer_t) * (size_t)n * sketch_size, cudaMemcpyDeviceToHost));
  SK_CHECK(cudaMemcpy(out_count, d_cnt, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
done:
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(d_seq);
  cudaFree(d_frags);
  cudaFree(d_out);
  cudaFree(d_cnt);
#else
  (void)device;
  std::vector<unsigned char> sm(smem + 64);
  memset(out, 0, sizeof(wfb_minmer_t) * (size_t)n * sketch_size);
  for (int b = 0; b < 1; ++b)
    wfb_sketch_kernel(b, 1, (const uint8_t*)seq_base, frags, n, kmer_size, sketch_size, npow2, out, out_count, sm.data());
#endif
  return rc;
}
