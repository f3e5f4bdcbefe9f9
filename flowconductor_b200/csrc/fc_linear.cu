// fc_linear.cu — the conditioner's dense layers on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), fp32-faithful.
//
// Replaces nn.Linear / MaskedLinear.forward of the conditioner networks
// (flowcon/nn/nets/resnet.py:39-56,90-99, flowcon/transforms/made.py:71-72,266-272) and — for the final layer —
// the element-wise bijection that consumes its output (rational_quadratic.py:13-181, coupling.py:279-293,549-582):
// the spline runs in the GEMM epilogue straight out of tensor memory, so the [B, D_t*P] parameter tensor
// (2.9 GB per layer at cfg 2) never exists in HBM.
//
// Precision: the reference computes these GEMMs in fp32 (no TF32).  One tf32 UMMA per product would lose 13
// mantissa bits, so every operand is split  a = a_hi + a_lo  (both tf32-representable, |a - a_hi - a_lo| <=
// 2^-22 |a|) and the product is accumulated as  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  in the fp32 TMEM accumulator
// ("3xTF32": error ~2^-21 per product, the same size as the summation-order noise of an fp32 GEMM with K = 256).
//   * weights: split once by fc_linear_pack into a [hi | lo] pair of K-major planes (also folds the MADE mask,
//     the per-feature padding P -> P_pad and the coupling column scatter into the layout);
//   * activations: TMA lands the raw fp32 tile in shared memory, two converter warps rewrite it in place as a_hi
//     (optionally after ReLU — the residual blocks are pre-activation, resnet.py:41-47) and write a_lo next to it.
//
// Accumulation: the tensor core truncates (rounds toward zero) every time an MMA result is added to the fp32
// accumulator in TMEM; over the K/8 = 32 dependent additions of a K = 256 dot product that is a systematic error
// of ~1.5e-6 relative, 5x the rounding noise of an fp32 FMA chain.  The accumulator is therefore drained every
// `chunk` pipeline stages (default 32 k-values = 4 UMMA k-steps): the MMA warp starts a fresh partial sum in the
// other TMEM buffer while the epilogue warps add the finished partial into fp32 REGISTER accumulators with
// round-to-nearest.  The partials are 8x smaller and 8x shorter, which brings the GEMM error down to that of the
// fp32 cuBLAS GEMM the reference runs (measured on K = 256 Gaussian operands, scripts/check_linear.py: rms error
// 1.8e-6 undrained, 4.6e-7 at 64, 2.5e-7 at 32; cuBLAS fp32 2.9e-7).
//
// Kernel: persistent, one CTA per SM, 16 warps (register budget re-balanced with setmaxnreg):
//   warp 0      TMA producer   : per stage one box of A (128 x BK) and the hi/lo boxes of W (BN x BK each)
//   warp 1      UMMA issuer    : 3 x BK/8 tcgen05.mma (128 x BN x 8, kind::tf32) per stage, commit -> barriers
//   warp 2      TMEM allocator
//   warps 4-7   converters     : raw A tile -> (a_hi, a_lo)
//   warps 8-15  epilogue       : tcgen05.ld each partial accumulator (double-buffered in TMEM, so draining chunk i
//                                overlaps the MMAs of chunk i+1) into registers, then + bias and either
//                                ReLU/residual/store or the rational-quadratic spline of 8 (K=8) / 4 (K=16)
//                                features per 192-column tile.
#include "fc_common.cuh"
#include "fc_tc.cuh"

namespace fc {

using namespace tc;

constexpr int kLinThreads = 512;
constexpr int kBM = 128;
constexpr int kConvWarp0 = 4;
constexpr int kEpiWarp0 = 8;
constexpr int kNumConv = 128;  // converter threads (warps 4-7)
constexpr int kRegsProducer = 40, kRegsConverter = 56, kRegsEpilogue = 208;  // 128*(40+56) + 256*208 == 65536

struct LinArgs {
  int M;
  int num_k_stages;  // pipeline stages (BK k-values each) per output tile
  int chunk;         // stages per partial accumulator
  int num_m_tiles, num_n_tiles;
  int n_pad;         // rows of one weight plane (lo plane starts at row n_pad of the weight tensor map)
  int relu_in;       // ReLU applied to A while splitting
  const float* bias;  // [n_pad]
};

struct StoreEpi {
  float* out;
  int64_t ldo;
  const float* residual;
  int64_t ldr;
  int n_out;
  int relu_out;
};

struct RqsEpi {
  const float* x;
  int64_t ldx;
  float* y;
  int64_t ldy;
  float* lad;
  int accumulate;
  const int32_t* tcols;
  const int32_t* ccols;
  int n_copy;
  int D_t;
  RqsParams c;
  int32_t* status;
};

template <int BN, int BK, int STAGES>
struct LinSmem {
  static constexpr int A_BYTES = kBM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_BYTES = 8 * (3 * STAGES + 4) + 16;
  static constexpr int LAD_BYTES = 2 * kBM * 4;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + LAD_BYTES + 1024;  // + alignment slack
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles must keep 1024-byte alignment");
};

template <int N>
__device__ __forceinline__ void set_max_regs_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void set_max_regs_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// acc[0..N) (=|+=) N consecutive TMEM columns of this thread's lane.  16 columns per tcgen05.ld, software-pipelined:
// the additions of group j run while the load of group j+1 is in flight.
template <int N, bool kFirst>
__device__ __forceinline__ void drain_partial(uint32_t taddr, float* acc) {
  static_assert(N % 32 == 0, "column count per thread must be a multiple of 32");
  if (kFirst) {
#pragma unroll
    for (int j = 0; j < N; j += 32) tmem_ld32(taddr + (uint32_t)j, reinterpret_cast<uint32_t*>(acc + j));
    tmem_wait_ld();
    return;
  }
  uint32_t v[2][16];
  tmem_ld16(taddr, v[0]);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < N; j += 16) {
    const int cur = (j >> 4) & 1;
    if (j + 16 < N) tmem_ld16(taddr + (uint32_t)(j + 16), v[cur ^ 1]);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[j + i] += __uint_as_float(v[cur][i]);
    if (j + 16 < N) tmem_wait_ld();
  }
}

// EPI: 0 = store (bias, optional residual / ReLU), 1 = rational-quadratic spline with KC bins.
// RQS tile geometry: FEATS features of PPAD accumulator columns each (BN = FEATS * PPAD).
template <int EPI, int BN, int BK, int STAGES, int KC, int PPAD>
__global__ void __launch_bounds__(kLinThreads, 1)
    linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const LinArgs la, const StoreEpi se, const RqsEpi re) {
  using SM = LinSmem<BN, BK, STAGES>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_s = s32(smem_raw);
  const uint32_t base = (raw_s + 1023u) & ~1023u;
  unsigned char* const gbase = smem_raw + (base - raw_s);
  const uint32_t bars = base + STAGES * SM::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto conv_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (3 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);
  volatile uint32_t* const tmem_slot_g =
      reinterpret_cast<volatile uint32_t*>(gbase + STAGES * SM::STAGE_BYTES + 8 * (3 * STAGES + 4));
  float* const lad_x = reinterpret_cast<float*>(gbase + STAGES * SM::STAGE_BYTES + SM::BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;
  static_assert(2 * BN <= 512, "two partial accumulators must fit in tensor memory");

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), kNumConv);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;

  const int nk = la.num_k_stages;
  const int n_tiles = la.num_n_tiles;
  const int chunk = la.chunk;
  const int n_chunks = (nk + chunk - 1) / chunk;

  if (warp < kConvWarp0) {
    set_max_regs_dec<kRegsProducer>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int mt = blockIdx.x; mt < la.num_m_tiles; mt += gridDim.x) {
          for (int nt = 0; nt < n_tiles; ++nt) {
            for (int kc = 0; kc < nk; ++kc) {
              mbar_wait(empty_bar(s), ph ^ 1u);
              const uint32_t st = base + s * SM::STAGE_BYTES;
              mbar_expect_tx(full_bar(s), SM::A_BYTES + 2 * SM::B_BYTES);
              tma_load_2d(st, &tmA, kc * BK, mt * kBM, full_bar(s));
              tma_load_2d(st + 2 * SM::A_BYTES, &tmB, kc * BK, nt * BN, full_bar(s));
              tma_load_2d(st + 2 * SM::A_BYTES + SM::B_BYTES, &tmB, kc * BK, la.n_pad + nt * BN, full_bar(s));
              if (++s == STAGES) {
                s = 0;
                ph ^= 1u;
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- UMMA issuer
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_tf32(kBM, BN);
        int s = 0, acc = 0;
        uint32_t ph = 0, aph = 0;
        for (int mt = blockIdx.x; mt < la.num_m_tiles; mt += gridDim.x) {
          for (int nt = 0; nt < n_tiles; ++nt) {
            for (int k0 = 0; k0 < nk; k0 += chunk) {
              const int k1 = k0 + chunk < nk ? k0 + chunk : nk;
              mbar_wait(tempty_bar(acc), aph ^ 1u);
              tc_fence_after();
              const uint32_t d = tmem_base + (uint32_t)(acc * BN);
              for (int kc = k0; kc < k1; ++kc) {
                mbar_wait(full_bar(s), ph);
                mbar_wait(conv_bar(s), ph);
                tc_fence_after();
                const uint32_t st = base + s * SM::STAGE_BYTES;
                const uint64_t a_hi = make_smem_desc(st, BK * 4);
                const uint64_t a_lo = make_smem_desc(st + SM::A_BYTES, BK * 4);
                const uint64_t b_hi = make_smem_desc(st + 2 * SM::A_BYTES, BK * 4);
                const uint64_t b_lo = make_smem_desc(st + 2 * SM::A_BYTES + SM::B_BYTES, BK * 4);
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                  const uint64_t o = (uint64_t)(kk * 2);  // 8 tf32 = 32 bytes = 2 x 16 B along K in the swizzle span
                  // small terms first, so they are not absorbed by the large one before they have been summed
                  umma_tf32_ss(d, a_lo + o, b_hi + o, idesc, (kc > k0 || kk > 0) ? 1u : 0u);
                  umma_tf32_ss(d, a_hi + o, b_lo + o, idesc, 1u);
                  umma_tf32_ss(d, a_hi + o, b_hi + o, idesc, 1u);
                }
                umma_commit(empty_bar(s));
                if (++s == STAGES) {
                  s = 0;
                  ph ^= 1u;
                }
              }
              umma_commit(tfull_bar(acc));
              if (++acc == 2) {
                acc = 0;
                aph ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp < kEpiWarp0) {
    // ------------------------------------------------------------------ converters: raw fp32 -> (hi, lo) tf32
    set_max_regs_dec<kRegsConverter>();
    const int ct = threadIdx.x - kConvWarp0 * 32;
    int s = 0;
    uint32_t ph = 0;
    const bool relu = la.relu_in != 0;
    for (int mt = blockIdx.x; mt < la.num_m_tiles; mt += gridDim.x) {
      for (int it = 0; it < n_tiles * nk; ++it) {
        mbar_wait(full_bar(s), ph);
        const uint32_t st = base + s * SM::STAGE_BYTES;
#pragma unroll
        for (int v = ct; v < SM::A_BYTES / 16; v += kNumConv) {
          const uint32_t addr = st + (uint32_t)v * 16u;
          float4 f = lds128(addr);
          if (relu) {
            f.x = fmaxf(f.x, 0.f);
            f.y = fmaxf(f.y, 0.f);
            f.z = fmaxf(f.z, 0.f);
            f.w = fmaxf(f.w, 0.f);
          }
          const uint32_t h0 = to_tf32(f.x), h1 = to_tf32(f.y), h2 = to_tf32(f.z), h3 = to_tf32(f.w);
          const uint32_t l0 = to_tf32(f.x - __uint_as_float(h0)), l1 = to_tf32(f.y - __uint_as_float(h1)),
                         l2 = to_tf32(f.z - __uint_as_float(h2)), l3 = to_tf32(f.w - __uint_as_float(h3));
          sts128(addr, h0, h1, h2, h3);
          sts128(addr + SM::A_BYTES, l0, l1, l2, l3);
        }
        fence_proxy_async_smem();
        mbar_arrive(conv_bar(s));
        if (++s == STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    set_max_regs_inc<kRegsEpilogue>();
    constexpr int NCOL = BN / 2;             // accumulator columns (and registers) per thread
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;  // which half of the tile's columns / features
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    int acc = 0;
    uint32_t aph = 0;
    unsigned status = 0;
    int parity = 0;
    for (int mt = blockIdx.x; mt < la.num_m_tiles; mt += gridDim.x) {
      const int64_t row = (int64_t)mt * kBM + q * 32 + lane;
      const bool valid = row < la.M;
      float lad_acc = 0.f;
      for (int nt = 0; nt < n_tiles; ++nt) {
        float av[NCOL];
        for (int ch = 0; ch < n_chunks; ++ch) {
          mbar_wait(tfull_bar(acc), aph);
          tc_fence_after();
          const uint32_t tacc = tmem_base + lane_sel + (uint32_t)(acc * BN + half * NCOL);
          if (ch == 0) {
            drain_partial<NCOL, true>(tacc, av);
          } else {
            drain_partial<NCOL, false>(tacc, av);
          }
          tc_fence_before();  // partial accumulator fully read by this warp: hand it back to the MMA warp
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          if (++acc == 2) {
            acc = 0;
            aph ^= 1u;
          }
        }
        if (EPI == 0) {
          const int n0 = nt * BN + half * NCOL;
          const float4* b4 = reinterpret_cast<const float4*>(la.bias + n0);
#pragma unroll
          for (int j = 0; j < NCOL / 4; ++j) {
            const float4 b = __ldg(b4 + j);
            float4 o;
            o.x = av[4 * j + 0] + b.x;
            o.y = av[4 * j + 1] + b.y;
            o.z = av[4 * j + 2] + b.z;
            o.w = av[4 * j + 3] + b.w;
            const int n = n0 + 4 * j;
            if (valid && n < se.n_out) {
              if (se.residual) {
                const float4 r = __ldg(reinterpret_cast<const float4*>(se.residual + row * se.ldr + n));
                o.x += r.x;
                o.y += r.y;
                o.z += r.z;
                o.w += r.w;
              }
              if (se.relu_out) {
                o.x = fmaxf(o.x, 0.f);
                o.y = fmaxf(o.y, 0.f);
                o.z = fmaxf(o.z, 0.f);
                o.w = fmaxf(o.w, 0.f);
              }
              *reinterpret_cast<float4*>(se.out + row * se.ldo + n) = o;
            }
          }
        } else {
          constexpr int FEATS = BN / PPAD;
          constexpr int FH = FEATS / 2;
          static_assert(FH * PPAD == NCOL, "feature groups must tile the column half exactly");
          if (nt == 0 && half == 0 && valid && re.n_copy > 0 && re.y != re.x) {
            for (int i = 0; i < re.n_copy; ++i) {  // identity columns (coupling.py:96-98)
              const int cc = __ldg(re.ccols + i);
              re.y[row * re.ldy + cc] = __ldg(re.x + row * re.ldx + cc);
            }
          }
#pragma unroll
          for (int f = 0; f < FH; ++f) {
            const int fl = half * FH + f;    // feature within the tile
            const int fg = nt * FEATS + fl;  // feature of the layer
            if (fg < re.D_t) {
              const float* bp = la.bias + (nt * BN + fl * PPAD);
              float* p = av + f * PPAD;
#pragma unroll
              for (int i = 0; i < PPAD; i += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bp + i));
                p[i + 0] += b.x;
                p[i + 1] += b.y;
                p[i + 2] += b.z;
                p[i + 3] += b.w;
              }
              const int col = re.tcols ? __ldg(re.tcols + fg) : fg;
              const float xv = valid ? __ldg(re.x + row * re.ldx + col) : 0.f;
              float yv, lv;
              rqs_eval<KC, true>(re.c, xv, p, yv, lv, status);
              if (valid) re.y[row * re.ldy + col] = yv;
              lad_acc += lv;
            }
          }
        }
      }
      if (EPI == 1) {
        // per-sample log|det J| (sum_except_batch, utils/torchutils.py:25-30): this thread summed its features in
        // order; the two column halves of a row are combined in a fixed order through shared memory
        float* ex = lad_x + parity * kBM;
        if (half == 1) ex[q * 32 + lane] = lad_acc;
        named_barrier_sync(1, 256);
        if (half == 0 && valid) {
          const float tot = lad_acc + ex[q * 32 + lane];
          re.lad[row] = re.accumulate ? re.lad[row] + tot : tot;
        }
        parity ^= 1;
      }
    }
    if (EPI == 1 && status != 0 && re.status) atomicOr(re.status, (int)status);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// [rows, cols] fp32 row-major matrix with row stride `ld` elements; box = box_rows x bk, swizzle span = bk * 4 bytes
static int make_map(CUtensorMap* m, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, int box_rows, int bk) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return FC_ERR_CUDA;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FC_OK : FC_ERR_CUDA;
}

// k-values per partial accumulator (see "Accumulation" at the top).  Short dot products leave less room between
// the tensor core's truncation error and the (smaller) rounding noise of an fp32 FMA chain of the same length, so
// they are drained after every stage.  FC_LINEAR_CHUNK_K overrides for experiments.
static int chunk_k(int K) {
  static int forced = [] {
    const char* e = getenv("FC_LINEAR_CHUNK_K");
    return e ? atoi(e) : 0;
  }();
  if (forced > 0) return forced < 16 ? 16 : forced;
  return K <= 64 ? 16 : 32;
}

template <int EPI, int BN, int BK, int STAGES, int KC, int PPAD>
static int launch_linear(const float* A, int64_t lda, int64_t M, int K, const fc_linear_weights* w, LinArgs la,
                         const StoreEpi& se, const RqsEpi& re, cudaStream_t stream) {
  using SM = LinSmem<BN, BK, STAGES>;
  if (w->n_pad % BN != 0 || w->k_pad % 32 != 0) return FC_ERR_INVALID_ARGUMENT;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, BK);
  if (rc != FC_OK) return rc;
  rc = make_map(&tmB, w->w, (uint64_t)2 * w->n_pad, (uint64_t)w->k_pad, (uint64_t)w->k_pad, BN, BK);
  if (rc != FC_OK) return rc;
  la.M = (int)M;
  la.num_k_stages = (K + BK - 1) / BK;
  la.chunk = chunk_k(K) / BK > 0 ? chunk_k(K) / BK : 1;
  la.num_m_tiles = (int)((M + kBM - 1) / kBM);
  la.n_pad = w->n_pad;
  la.bias = w->bias;
  auto kern = linear_tf32x3_kernel<EPI, BN, BK, STAGES, KC, PPAD>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL) != cudaSuccess)
      return FC_ERR_CUDA;
    configured = true;
  }
  const int grid = la.num_m_tiles < device_info().sm_count ? la.num_m_tiles : device_info().sm_count;
  kern<<<grid, kLinThreads, SM::TOTAL, stream>>>(tmA, tmB, la, se, re);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

static int check_operand(const float* A, int64_t lda, int64_t M, int K, const fc_linear_weights* w) {
  if (!A || !w || !w->w || !w->bias || M < 0 || K <= 0) return FC_ERR_INVALID_ARGUMENT;
  if (M >= (int64_t)1 << 31) return FC_ERR_UNSUPPORTED;
  if (K > w->k_pad) return FC_ERR_INVALID_ARGUMENT;
  // TMA: 16-byte aligned base and row pitch
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (lda & 3) || lda < K) return FC_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(w->w) & 15) || (reinterpret_cast<uintptr_t>(w->bias) & 15)) return FC_ERR_UNSUPPORTED;
  return FC_OK;
}

__global__ void pack_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ mask, int64_t ldm,
                            const float* __restrict__ bias, const int32_t* __restrict__ row_map,
                            const int32_t* __restrict__ col_map, int N, int K, int n_pad, int k_pad,
                            float* __restrict__ out_w, float* __restrict__ out_b) {
  const int64_t total = (int64_t)N * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / K), k = (int)(i % K);
    float wv = W[n * ldw + k];
    if (mask) wv = wv * mask[n * ldm + k];  // MaskedLinear: weight * mask (made.py:72)
    const int rn = row_map ? row_map[n] : n;
    const int ck = col_map ? col_map[k] : k;
    const uint32_t hi = tc::to_tf32(wv);
    const uint32_t lo = tc::to_tf32(wv - __uint_as_float(hi));
    out_w[(int64_t)rn * k_pad + ck] = __uint_as_float(hi);
    out_w[((int64_t)n_pad + rn) * k_pad + ck] = __uint_as_float(lo);
    if (k == 0) out_b[rn] = bias ? bias[n] : 0.f;
  }
}

}  // namespace fc

using namespace fc;

extern "C" int fc_linear_pack(const float* W, int64_t w_row_stride, const float* mask, int64_t mask_row_stride,
                              const float* bias, int32_t N, int32_t K, const int32_t* row_map, const int32_t* col_map,
                              int32_t n_pad, int32_t k_pad, float* w_packed, float* bias_packed, void* stream) {
  if (!W || !w_packed || !bias_packed || N <= 0 || K <= 0 || n_pad < N || k_pad < K) return FC_ERR_INVALID_ARGUMENT;
  if (k_pad % 32 != 0 || n_pad % 16 != 0) return FC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(w_packed, 0, sizeof(float) * 2 * (size_t)n_pad * k_pad, st) != cudaSuccess) return FC_ERR_CUDA;
  if (cudaMemsetAsync(bias_packed, 0, sizeof(float) * (size_t)n_pad, st) != cudaSuccess) return FC_ERR_CUDA;
  const int64_t total = (int64_t)N * K;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_kernel<<<blocks, 256, 0, st>>>(W, w_row_stride, mask, mask_row_stride, bias, row_map, col_map, N, K, n_pad, k_pad,
                                      w_packed, bias_packed);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

extern "C" int fc_linear_apply(const float* A, int64_t lda, int64_t M, int32_t K, const fc_linear_weights* w,
                               int32_t relu_in, float* out, int64_t ldo, int32_t n_out, int32_t relu_out,
                               const float* residual, int64_t ldr, void* stream) {
  int rc = check_operand(A, lda, M, K, w);
  if (rc != FC_OK) return rc;
  if (!out || n_out <= 0 || n_out > w->n_pad) return FC_ERR_INVALID_ARGUMENT;
  if ((n_out & 3) || (ldo & 3) || (reinterpret_cast<uintptr_t>(out) & 15)) return FC_ERR_UNSUPPORTED;
  if (residual && ((ldr & 3) || (reinterpret_cast<uintptr_t>(residual) & 15))) return FC_ERR_UNSUPPORTED;
  if (M == 0) return FC_OK;
  LinArgs la{};
  la.relu_in = relu_in;
  StoreEpi se{out, ldo, residual, ldr, n_out, relu_out};
  RqsEpi re{};
  constexpr int BN = 256;
  if (w->n_pad % BN != 0) return FC_ERR_INVALID_ARGUMENT;
  la.num_n_tiles = (n_out + BN - 1) / BN;
  return launch_linear<0, BN, 16, 4, 0, 32>(A, lda, M, K, w, la, se, re, (cudaStream_t)stream);
}

extern "C" int fc_linear_rqs_apply(const float* hidden, int64_t ldh, int64_t B, int32_t H, const fc_linear_weights* w,
                                   int32_t relu_in, const float* x, int64_t x_row_stride, float* y,
                                   int64_t y_row_stride, float* logabsdet, int32_t accumulate_logabsdet, int32_t D_t,
                                   fc_cols tcols, fc_cols ccols, const fc_rqs_config* cfg, int32_t* status,
                                   void* stream) {
  RqsParams c;
  int rc = make_rqs_params(cfg, c);
  if (rc != FC_OK) return rc;
  rc = check_operand(hidden, ldh, B, H, w);
  if (rc != FC_OK) return rc;
  if (!x || !y || !logabsdet || D_t <= 0) return FC_ERR_INVALID_ARGUMENT;
  if (tcols.idx && tcols.n != D_t) return FC_ERR_INVALID_ARGUMENT;
  if (c.tails != FC_TAILS_LINEAR) return FC_ERR_UNSUPPORTED;  // fused epilogue: linear tails (P = 3K-1) only
  if (B == 0) return FC_OK;
  LinArgs la{};
  la.relu_in = relu_in;
  StoreEpi se{};
  RqsEpi re{x, x_row_stride, y, y_row_stride, logabsdet, accumulate_logabsdet, tcols.idx, ccols.idx, ccols.n, D_t, c, status};
  constexpr int BN = 192;
  if (c.K == 8) {
    constexpr int PPAD = 24;
    la.num_n_tiles = (D_t + BN / PPAD - 1) / (BN / PPAD);
    if (w->n_pad < la.num_n_tiles * BN) return FC_ERR_INVALID_ARGUMENT;
    return launch_linear<1, BN, 16, 5, 8, PPAD>(hidden, ldh, B, H, w, la, se, re, (cudaStream_t)stream);
  }
  if (c.K == 16) {
    constexpr int PPAD = 48;
    la.num_n_tiles = (D_t + BN / PPAD - 1) / (BN / PPAD);
    if (w->n_pad < la.num_n_tiles * BN) return FC_ERR_INVALID_ARGUMENT;
    return launch_linear<1, BN, 16, 5, 16, PPAD>(hidden, ldh, B, H, w, la, se, re, (cudaStream_t)stream);
  }
  return FC_ERR_UNSUPPORTED;
}
