// tcgen05 implicit-GEMM convolution for sm_100a (stride 1, arbitrary tap set, up to 3 virtually
// concatenated NHWC bf16 sources, fused bias + activation + residual epilogue).
//
// GEMM view:  D[128 pixels, BLOCK_N couts] += sum over (tap, source, 64-channel chunk) of
//             A[128 pixels, 64 ch] * B[BLOCK_N couts, 64 ch]^T        (bf16 x bf16 -> fp32 in TMEM)
//   * A tile = one TMA 4-D box {64 ch, tw, th, 1} of the NHWC source at pixel offset (dx,dy) of the
//     tap: the box lands in shared memory as 128 rows x 128 B, K-major, 128-B swizzled -- exactly
//     the canonical UMMA operand layout, so there is no im2col pass; out-of-image rows/columns
//     (the convolution's zero padding) and channels beyond the source's C are zero-filled by TMA.
//   * B tile = TMA 3-D box {64 k, BLOCK_N, 1 tap} of the packed weights [tap][Cout_pad][Ktot].
//   * warp 0 = TMA producer, warp 1 = MMA issuer (single thread issues tcgen05.mma, accumulators
//     double-buffered in TMEM), warps 2-9 = epilogue (tcgen05.ld -> bias/act/residual -> global).
//   * persistent CTAs, static round-robin tile schedule, mbarrier full/empty smem ring.
// Restates torch.nn.Conv2d(stride=1) + LeakyReLU/ELU/ReLU as used at lite_flow_net.py:98-240 and
// depth_decoder.py / torchvision BasicBlock (BN folded by the weight packer).
#include "tc_ptx.cuh"

#ifndef DFVO_HOSTSIM
#include <cuda.h>
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

namespace dfvo {

void conv_tc_tile_shape(int H, int W, int* tw, int* th) {
  // 128 pixels per tile as tw x th (powers of two); minimise padded area, prefer wide tiles.
  long long best = -1;
  int bw = 128, bh = 1;
  for (int w = 128; w >= 1; w >>= 1) {
    int h = 128 / w;
    long long area = (long long)cdiv(W, w) * w * (long long)cdiv(H, h) * h;
    if (best < 0 || area < best) { best = area; bw = w; bh = h; }
  }
  *tw = bw; *th = bh;
}

struct ConvTcK {
  int N, H, W, tw, th, tiles_x, tiles_y, n_blocks, ntiles;
  int nsrc, srcC[3];
  int ntaps;
  int8_t dy[49], dx[49];        // stride 1: input offsets; stride 2: offsets in units of double-pixels (floor((k-pad)/2))
  int8_t py[49], px[49];        // stride 2: pixel phase of the tap inside the 2x2 cell
  int stride, src_pitch;        // src_pitch (elements) = channel offset of the odd pixel inside a double-pixel
  int block_n, stages, acc_stride, tmem_cols;
  int Cout, Cout_pad, act, out_f32, zero_pad_to;
  int chunk, esize, round_tf32;   // channels per 128-byte operand row (64 bf16 / 32 tf32), operand element size
  const float* bias;
  void* out; long long oN, oH, oW;
  const void* res; long long rN, rH, rW;
};

#ifndef DFVO_HOSTSIM
// =============================================================================================
//                                       device side
// =============================================================================================
#define TC_THREADS 320
#define TC_A_BYTES 16384

template <int TF32>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
          const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ ConvTcK p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B needs 1024-B alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t stage_bytes = TC_A_BYTES + (uint32_t)p.block_n * 128u;
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;    // 8-byte aligned
  // barrier layout: full[stages], empty[stages], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * p.stages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (size_t)p.stages * stage_bytes + 8u * (2 * p.stages + 4));
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_trigger();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0); prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                  // from here on: activations of the previous kernel / our output buffers

  int chunks_total = 0;
  for (int s = 0; s < p.nsrc; ++s) chunks_total += (p.srcC[s] + p.chunk - 1) / p.chunk;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // whole warp walks the loop, one elected lane issues (see tc_ptx.cuh::elect_one)
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int n = t % p.N; const int nb = t / p.N;
      const int x0 = tx * p.tw, y0 = ty * p.th;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        int kofs = 0;
        for (int s = 0; s < p.nsrc; ++s) {
          const CUtensorMap* tm = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : &tmA2);
          for (int c0 = 0; c0 < p.srcC[s]; c0 += p.chunk) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (elect_one()) {
              const uint32_t sa = base + (uint32_t)stage * stage_bytes;
              mbar_expect_tx(full_bar(stage), stage_bytes);
              if (p.stride == 2)
                tma_load_5d(sa, tm, full_bar(stage), c0 + p.px[tap] * p.src_pitch, x0 + p.dx[tap], p.py[tap], y0 + p.dy[tap], n);
              else
                tma_load_4d(sa, tm, full_bar(stage), c0, x0 + p.dx[tap], y0 + p.dy[tap], n);
              tma_load_3d(sa + TC_A_BYTES, &tmB, full_bar(stage), kofs + c0, nb * p.block_n, tap);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          kofs += p.srcC[s];
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =========================================
    const uint32_t idesc = tc_idesc(TF32, p.block_n);
    const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024 B, version 1, SWIZZLE_128B
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.acc_stride);
      uint32_t fresh = 0;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        for (int s = 0; s < p.nsrc; ++s) {
          for (int c0 = 0; c0 < p.srcC[s]; c0 += p.chunk) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = base + (uint32_t)stage * stage_bytes;
              const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (1u << 16), b_lo = (((sa + TC_A_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
              const int rem = p.srcC[s] - c0;
              const int nks = ((rem >= p.chunk ? p.chunk : rem) * p.esize) >> 5;       // 32-byte K steps with real channels
              if (nks == 4) {
                tc_mma_lohi<TF32>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, fresh);
                tc_mma_lohi<TF32>(tmem_d, a_lo + 2u, d_hi, b_lo + 2u, d_hi, idesc, 1u);   // +32 B inside the 128-B swizzle atom
                tc_mma_lohi<TF32>(tmem_d, a_lo + 4u, d_hi, b_lo + 4u, d_hi, idesc, 1u);
                tc_mma_lohi<TF32>(tmem_d, a_lo + 6u, d_hi, b_lo + 6u, d_hi, idesc, 1u);
              } else {
                for (int ks = 0; ks < nks; ++ks)
                  tc_mma_lohi<TF32>(tmem_d, a_lo + 2u * ks, d_hi, b_lo + 2u * ks, d_hi, idesc, ks == 0 ? fresh : 1u);
              }
              tc_commit(empty_bar(stage));
            }
            __syncwarp();
            fresh = 1u;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (elect_one()) tc_commit(tfull_bar(acc));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ===================================== epilogue warps ====================================
    // 8 warps: TMEM lane quadrant = warp % 4 (hardware rule), two warps per quadrant split the columns.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int nchunks = p.block_n >> 4;
    const int ch_begin = (ew < 4) ? 0 : ((nchunks + 1) >> 1);
    const int ch_end = (ew < 4) ? ((nchunks + 1) >> 1) : nchunks;
    // bias staging belongs to the epilogue warps alone (they idle until the first accumulator is ready anyway), so the
    // producers and the MMA warp start on the barrier-init sync instead of waiting for a global load: ~1 us off the
    // critical path of each of the ~115 launches per frame.
    for (int i = threadIdx.x - 64; i < p.Cout_pad; i += 256) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float4* bias4 = reinterpret_cast<const float4*>(bias_s);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int n = t % p.N; const int nb = t / p.N;
      const int x = tx * p.tw + (row % p.tw), y = ty * p.th + (row / p.tw);
      const bool inb = x < p.W && y < p.H;
      const int cbase = nb * p.block_n;
      const long long opix = n * p.oN + y * p.oH + x * p.oW;
      const long long rpix = n * p.rN + y * p.rH + x * p.rW;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);

      TcEpi ep; ep.Cout = p.Cout; ep.zero_pad_to = p.zero_pad_to; ep.act = p.act; ep.out_f32 = p.out_f32; ep.round_tf32 = p.round_tf32; ep.out = p.out; ep.res = p.res;
      // lean path (tc_ptx.cuh::tc_epilogue16_fast): bf16 output, no residual, LeakyReLU / ReLU / identity, aligned pixels
      const bool fast_launch = !p.out_f32 && p.res == nullptr && p.act <= ACT_RELU && (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0 &&
                               (p.oW & 7) == 0 && (p.oH & 7) == 0 && (p.oN & 7) == 0;
      const float slope = p.act == ACT_LEAKY ? 0.1f : (p.act == ACT_RELU ? 0.f : 1.f);
      auto process = [&](const uint32_t* v, int col) {
        const int c = cbase + col;
        if (fast_launch && c + 16 <= p.Cout) tc_epilogue16_fast(v, bias_s + c, slope, reinterpret_cast<__nv_bfloat16*>(p.out) + opix + c, inb);
        else if (inb && c < p.zero_pad_to) tc_epilogue16_call(ep, v, bias4, c, opix, rpix);
      };

      for (int ch = ch_begin; ch < ch_end; ch += 2) {
        uint32_t v0[16], v1[16];
        const bool two = ch + 1 < ch_end;
        __syncwarp();                              // tcgen05.ld is .sync.aligned: reconverge first
        tc_ld16_nowait(taddr0 + (uint32_t)(ch * 16), v0);
        if (two) tc_ld16_nowait(taddr0 + (uint32_t)(ch * 16 + 16), v1);
        tc_ld_wait16(v0);
        if (two) tc_ld_wait16(v1);
        process(v0, ch * 16);
        if (two) process(v1, ch * 16 + 16);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// =============================================================================================
//                                        host side
// =============================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int tc_encode_map(void* map, const void* ptr, int rank, const unsigned long long* dims_, const unsigned long long* strides_, const unsigned* box_,
                  int esize, int swizzle_bytes) {
  cuuint64_t dims[5], str[5];
  cuuint32_t box[5];
  for (int i = 0; i < rank; ++i) { dims[i] = dims_[i]; box[i] = box_[i]; if (i < rank - 1) str[i] = strides_[i]; }
  CUtensorMap* m = reinterpret_cast<CUtensorMap*>(map);
  const cuuint64_t* strides_bytes = str;
  PFN_encodeTiled enc = get_encode();
  DFVO_REQUIRE(enc != nullptr, DFVO_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFVO_REQUIRE(r == CUDA_SUCCESS, DFVO_ECUDA, "cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu %llu box %u %u %u)", (int)r,
               rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1], box[2]);
  return DFVO_OK;
}

struct ConvTcPlanImpl {
  CUtensorMap tmA[3], tmB;
  ConvTcK k;
  int grid;
  size_t smem;
};

static int g_num_sms = 0;
void tc_launch_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int grid, int block, size_t smem, cudaStream_t s) {
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("DFVO_PDL"); pdl = !(e && atoi(e) == 0); }
  memset(cfg, 0, sizeof(*cfg));
  cfg->gridDim = dim3(grid); cfg->blockDim = dim3(block); cfg->dynamicSmemBytes = smem; cfg->stream = s;
  attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr->val.programmaticStreamSerializationAllowed = 1;
  cfg->attrs = attr; cfg->numAttrs = pdl ? 1 : 0;
}
int tc_num_sms() {
  if (!g_num_sms) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}
static int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                      int esize) {
  unsigned long long d[5], st[5]; unsigned b[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; if (i < rank - 1) st[i] = strides_bytes[i]; }
  return tc_encode_map(m, ptr, rank, d, st, b, esize, 128);
}

static int build_plan(const ConvTc& c, ConvTcPlanImpl* pl) {
  DFVO_REQUIRE(c.nsrc >= 1 && c.nsrc <= 3 && c.ntaps >= 1 && c.ntaps <= 49, DFVO_EINVAL, "conv_tc: nsrc/ntaps");
  DFVO_REQUIRE(c.Cout_pad % 16 == 0 && c.Cout_pad >= 16, DFVO_EINVAL, "conv_tc: Cout_pad %d must be a multiple of 16", c.Cout_pad);
  ConvTcK& k = pl->k;
  memset(&k, 0, sizeof(k));
  k.N = c.N; k.H = c.H; k.W = c.W;
  conv_tc_tile_shape(c.H, c.W, &k.tw, &k.th);
  k.tiles_x = cdiv(c.W, k.tw); k.tiles_y = cdiv(c.H, k.th);
  // N tile: whole Cout_pad if <= 256 else 128-wide blocks
  k.block_n = c.Cout_pad <= 256 ? c.Cout_pad : 128;
  DFVO_REQUIRE(c.Cout_pad % k.block_n == 0, DFVO_EINVAL, "conv_tc: Cout_pad %d not divisible by block_n %d", c.Cout_pad, k.block_n);
  k.n_blocks = c.Cout_pad / k.block_n;
  k.ntiles = k.tiles_x * k.tiles_y * c.N * k.n_blocks;
  k.nsrc = c.nsrc; k.ntaps = c.ntaps;
  const int es = c.esize == 4 ? 4 : 2;
  k.esize = es; k.chunk = 128 / es; k.round_tf32 = c.round_out_tf32;
  int ktot = 0;
  for (int s = 0; s < c.nsrc; ++s) {
    DFVO_REQUIRE(c.src[s].C % 16 == 0 && c.src[s].C > 0, DFVO_EINVAL, "conv_tc: source %d channels %d not a multiple of 16", s, c.src[s].C);
    DFVO_REQUIRE(((uintptr_t)c.src[s].p & 15) == 0 && (c.src[s].sW * es) % 16 == 0 && (c.src[s].sH * es) % 16 == 0 && (c.src[s].sN * es) % 16 == 0,
                 DFVO_EINVAL, "conv_tc: source %d not 16-byte aligned/strided", s);
    k.srcC[s] = c.src[s].C; ktot += c.src[s].C;
  }
  k.stride = c.stride == 2 ? 2 : 1;
  if (k.stride == 2) {
    const int inH = c.inH > 0 ? c.inH : 2 * c.H, inW = c.inW > 0 ? c.inW : 2 * c.W;
    DFVO_REQUIRE(c.nsrc == 1 && inH % 2 == 0 && inW % 2 == 0 && inH == 2 * c.H && inW == 2 * c.W, DFVO_EINVAL,
                 "conv_tc stride 2: needs one source and even input size = 2x output");
    DFVO_REQUIRE(c.src[0].sH == (long long)inW * c.src[0].sW, DFVO_EINVAL, "conv_tc stride 2: rows must be contiguous");
    k.src_pitch = (int)c.src[0].sW;
    for (int t = 0; t < c.ntaps; ++t) {
      const int oy = c.dy[t], ox = c.dx[t];                 // input-pixel offsets
      k.py[t] = (int8_t)(((oy % 2) + 2) % 2); k.dy[t] = (int8_t)((oy - k.py[t]) / 2);
      k.px[t] = (int8_t)(((ox % 2) + 2) % 2); k.dx[t] = (int8_t)((ox - k.px[t]) / 2);
    }
  } else {
    memcpy(k.dy, c.dy, sizeof(k.dy)); memcpy(k.dx, c.dx, sizeof(k.dx));
  }
  k.Cout = c.Cout; k.Cout_pad = c.Cout_pad; k.act = c.act; k.out_f32 = c.out_f32;
  k.zero_pad_to = c.zero_pad_to > c.Cout ? c.zero_pad_to : c.Cout;
  k.bias = c.bias; k.out = c.out; k.oN = c.oN; k.oH = c.oH; k.oW = c.oW;
  k.res = c.residual; k.rN = c.rN; k.rH = c.rH; k.rW = c.rW;
  // TMEM: two accumulators of block_n columns
  int acc_stride = 32; while (acc_stride < k.block_n) acc_stride <<= 1;
  k.acc_stride = acc_stride; k.tmem_cols = 2 * acc_stride;
  const size_t stage_bytes = TC_A_BYTES + (size_t)k.block_n * 128;
  const size_t fixed = 1024 /*align slack*/ + 8 * 64 /*barriers*/ + 64 + (size_t)c.Cout_pad * 4;
  int stages = (int)((200 * 1024 - fixed) / stage_bytes);
  if (stages > 8) stages = 8;
  DFVO_REQUIRE(stages >= 2, DFVO_EINVAL, "conv_tc: tile does not fit in shared memory");
  k.stages = stages;
  pl->smem = fixed + (size_t)stages * stage_bytes;
  pl->grid = k.ntiles < tc_num_sms() ? k.ntiles : tc_num_sms();
  // tensor maps
  for (int s = 0; s < 3; ++s) {
    const ConvTcSource& src = c.src[s < c.nsrc ? s : 0];
    if (k.stride == 2) {
      // view [N][H/2][2][W/2][pitch+C]: a double-pixel holds the even pixel's channels at [0,C) and the odd pixel's
      // at [pitch, pitch+C); (py, px) of a tap select the phase, the box walks W/2 x H/2 cells of the output tile
      const int inW2 = c.W, inH2 = c.H;
      cuuint64_t dims[5] = {(cuuint64_t)(src.sW + src.C), (cuuint64_t)inW2, 2, (cuuint64_t)inH2, (cuuint64_t)c.N};
      cuuint64_t str[4] = {(cuuint64_t)src.sW * 2 * es, (cuuint64_t)src.sH * es, (cuuint64_t)src.sH * 2 * es, (cuuint64_t)src.sN * es};
      cuuint32_t box[5] = {(cuuint32_t)k.chunk, (cuuint32_t)k.tw, 1, (cuuint32_t)k.th, 1};
      int rc = encode_map(&pl->tmA[s], src.p, 5, dims, str, box, es);
      if (rc) return rc;
      continue;
    }
    const int inW = c.inW > 0 ? c.inW : c.W, inH = c.inH > 0 ? c.inH : c.H;
    cuuint64_t dims[4] = {(cuuint64_t)src.C, (cuuint64_t)inW, (cuuint64_t)inH, (cuuint64_t)c.N};
    cuuint64_t str[3] = {(cuuint64_t)src.sW * es, (cuuint64_t)src.sH * es, (cuuint64_t)src.sN * es};
    cuuint32_t box[4] = {(cuuint32_t)k.chunk, (cuuint32_t)k.tw, (cuuint32_t)k.th, 1};
    int rc = encode_map(&pl->tmA[s], src.p, 4, dims, str, box, es);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)c.Cout_pad, (cuuint64_t)c.ntaps};
    cuuint64_t str[2] = {(cuuint64_t)ktot * es, (cuuint64_t)ktot * es * (cuuint64_t)c.Cout_pad};
    cuuint32_t box[3] = {(cuuint32_t)k.chunk, (cuuint32_t)k.block_n, 1};
    int rc = encode_map(&pl->tmB, c.w, 3, dims, str, box, es);
    if (rc) return rc;
  }
  return DFVO_OK;
}

std::atomic<long long> g_launch_count{0};
int g_tc_prof_on = 0;
static double g_prof_flops = 0.0;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static std::vector<std::string> g_prof_desc;      // per-launch layer description (DFVO_TC_TRACE=1 prints them with their times)

void conv_tc_profile_enable(int on) {
  g_tc_prof_on = on;
  if (on) {
    for (auto& e : g_prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    g_prof_events.clear();
    g_prof_desc.clear();
    g_prof_flops = 0.0;
  }
}

void conv_tc_profile_read(double* ms, long long* launches, double* flops) {
  double tot = 0.0;
  const bool trace = getenv("DFVO_TC_TRACE") != nullptr;
  for (size_t i = 0; i < g_prof_events.size(); ++i) {
    auto& e = g_prof_events[i];
    cudaEventSynchronize(e.second);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e.first, e.second) == cudaSuccess) tot += t;
    if (trace && i < g_prof_desc.size()) fprintf(stderr, "conv_tc %4zu %8.2f us  %s\n", i, t * 1e3, g_prof_desc[i].c_str());
  }
  *ms = tot; *launches = (long long)g_prof_events.size(); *flops = g_prof_flops;
}

bool tc_prof_begin(cudaStream_t s, TcProf* p) {
  if (!g_tc_prof_on) return false;
  cudaEventCreate(&p->e0); cudaEventCreate(&p->e1); cudaEventRecord(p->e0, s);
  return true;
}
void tc_prof_end(cudaStream_t s, const TcProf& p, double flops, const char* desc) {
  cudaEventRecord(p.e1, s);
  g_prof_events.push_back({p.e0, p.e1});
  g_prof_flops += flops;
  g_prof_desc.push_back(desc);
}

int conv_tc_single(const ConvTc& c, cudaStream_t s);
int conv_tc(const ConvTc& c, cudaStream_t s) {
  int rc = DFVO_OK;
  if (conv_chain_take(c, s, &rc)) return rc;              // deferred into the open layer chain (conv_chain.cu)
  if (rc) return rc;
  return conv_tc_single(c, s);
}

int conv_tc_single(const ConvTc& c, cudaStream_t s) {
  if (conv_halo_supported(c)) return conv_halo(c, s);
  ConvTcPlanImpl pl;
  int rc = build_plan(c, &pl);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    DFVO_CUDA(cudaFuncSetAttribute(k_conv_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DFVO_CUDA(cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  ++g_launch_count;
  TcProf pr;
  const bool prof = tc_prof_begin(s, &pr);
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr;
  tc_launch_config(&cfg, &attr, pl.grid, TC_THREADS, pl.smem, s);
  if (pl.k.esize == 4) DFVO_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<1>, pl.tmA[0], pl.tmA[1], pl.tmA[2], pl.tmB, pl.k));
  else DFVO_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<0>, pl.tmA[0], pl.tmA[1], pl.tmA[2], pl.tmB, pl.k));
  if (prof) {
    char d[256];
    snprintf(d, sizeof(d), "tap  N%d %dx%d s%d taps%d src[%d,%d,%d] cout%d/%d bn%d tile%dx%d stages%d grid%d tiles%d gflop %.3f", c.N, c.H, c.W,
             pl.k.stride, c.ntaps, c.src[0].C, c.nsrc > 1 ? c.src[1].C : 0, c.nsrc > 2 ? c.src[2].C : 0, c.Cout, c.Cout_pad, pl.k.block_n,
             pl.k.tw, pl.k.th, pl.k.stages, pl.grid, pl.k.ntiles, c.flops * 1e-9);
    tc_prof_end(s, pr, c.flops, d);
  }
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

#else  // DFVO_HOSTSIM ---------------------------------------------------------------------------
// CPU test build only: consumes the SAME ConvTc description and packed weights as the device
// kernel and applies the same TMA semantics (zero fill outside the image / beyond a source's C,
// virtual concat of sources, tap offsets), so the layer wiring and the weight packer can be
// validated without a GPU.  Never compiled into the product library.
std::atomic<long long> g_launch_count{0};
void conv_tc_profile_enable(int) {}
void conv_tc_profile_read(double* ms, long long* launches, double* flops) { *ms = 0; *launches = 0; *flops = 0; }
int conv_tc(const ConvTc& c, cudaStream_t) {
  int ktot = 0;
  for (int s = 0; s < c.nsrc; ++s) ktot += c.src[s].C;
  int zp = c.zero_pad_to > c.Cout ? c.zero_pad_to : c.Cout;
  std::vector<float> acc(c.Cout_pad);
  for (int n = 0; n < c.N; ++n)
    for (int y = 0; y < c.H; ++y)
      for (int x = 0; x < c.W; ++x) {
        for (int co = 0; co < c.Cout_pad; ++co) acc[co] = 0.f;
        for (int t = 0; t < c.ntaps; ++t) {
          const int st = c.stride == 2 ? 2 : 1;
          int iy = y * st + c.dy[t], ix = x * st + c.dx[t];
          const int inW = c.inW > 0 ? c.inW : c.W * st, inH = c.inH > 0 ? c.inH : c.H * st;
          if (iy < 0 || iy >= inH || ix < 0 || ix >= inW) continue;
          int kofs = 0;
          for (int s = 0; s < c.nsrc; ++s) {
            const size_t eoff = n * c.src[s].sN + iy * c.src[s].sH + ix * c.src[s].sW;
            for (int ci = 0; ci < c.src[s].C; ++ci) {
              const float av = c.esize == 4 ? ((const float*)c.src[s].p)[eoff + ci] : __bfloat162float(((const bf16*)c.src[s].p)[eoff + ci]);
              if (av == 0.f) continue;
              const size_t woff = ((size_t)t * c.Cout_pad) * ktot + kofs + ci;
              for (int co = 0; co < c.Cout_pad; ++co)
                acc[co] += av * (c.esize == 4 ? ((const float*)c.w)[woff + (size_t)co * ktot] : __bfloat162float(((const bf16*)c.w)[woff + (size_t)co * ktot]));
            }
            kofs += c.src[s].C;
          }
        }
        for (int co = 0; co < zp; ++co) {
          float v = 0.f;
          if (co < c.Cout) {
            v = acc[co] + (c.bias ? c.bias[co] : 0.f);
            if (c.residual) {
              if (c.out_f32) v += ((const float*)c.residual)[n * c.rN + y * c.rH + x * c.rW + co];
              else v += __bfloat162float(((const bf16*)c.residual)[n * c.rN + y * c.rH + x * c.rW + co]);
            }
            v = apply_act(v, c.act);
          }
          if (c.out_f32) ((float*)c.out)[n * c.oN + y * c.oH + x * c.oW + co] = v;
          else ((bf16*)c.out)[n * c.oN + y * c.oH + x * c.oW + co] = __float2bfloat16_rn(v);
        }
      }
  return DFVO_OK;
}
#endif

}  // namespace dfvo
