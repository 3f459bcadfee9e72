// Convolutions of the hot path as tcgen05 / TMEM implicit GEMMs over the zero-bordered
// "planes" layouts of planes.cuh (reference: sgtapose/lib/model/networks/dla.py:41-69
// BasicBlock, :157-175 Root, :212-216 project, :241-270 stems, :302-312 conv levels, :538-550
// DeformConv; base_model.py:121-135 heads).
//
//   y[p, o] = act( scale[o] * sum_k A[p, k] * Wt[k, o] + shift[o] (+ residual[p, o]) )
//
// Two kernels share the MMA issue loop and the epilogue:
//
//  conv_shift_kernel  (3x3 / 1x1, stride 1, Cin % 64 == 0: ~75 % of the network's FLOPs)
//      The A operand is never gathered.  Output rows m0..m0+127 of tap (dy,dx) read input rows
//      m0 + dy*(W+2) + dx .. +127, and a PL buffer already IS the SWIZZLE_128B operand image,
//      so per 64-channel chunk the TMA engine bulk-copies three row bands (one per dy, 144
//      rows) and the three dx taps are UMMA descriptors that start one row apart inside the
//      band.  Warp roles: 1 TMA thread, 1 MMA thread, 4 epilogue warps; persistent CTAs with a
//      double-buffered TMEM accumulator so the epilogue of tile i overlaps the MMAs of i+1.
//
//  conv_gather_kernel (A built by 8 producer warps, 8 lanes per 128-byte row so every
//      warp-wide load covers whole cache lines)
//      DCN   : A[p,(tap,c)] = sigmoid(mask) * bilinear(x[:,c], p + tap + offset); the zero
//              border of the frame implements "zero outside" (upstream DCNv2 writes the
//              [9*Cin, H*W] column matrix to HBM and reads it back for cuBLAS);
//      STRIDE: 3x3 stride-2 convolutions on PL inputs (one per DLA level);
//      SMALLC: convolutions on SC inputs (C in {4,16,32}: 7x7 stems, level0, level1, level2
//              entry), K blocks = 128-byte runs of consecutive pixels described by a table.
//
// NS = 1: bf16 planes, one MMA per K step.  NS = 2: fp16 hi/lo planes, three MMAs per K step
// into two accumulators (D0 = hi*hi, D1 = hi*lo + lo*hi), y = D0 + D1 * 2^-11.
#include "common.cuh"
#include "planes.cuh"
#include "umma.cuh"

namespace sgta {
using namespace umma;

constexpr int TM = 128;
constexpr int SMEM_LIMIT = 227 * 1024;
// producer warps of the gather kernel: the DCN producers are bound by dependent-instruction latency, not by issue
// slots or loads (tools/dcn_bench.py), so they get 16 warps (2 rows per lane) where the plain gathers keep 8 (4 rows;
// measured with 16: stem unchanged, level0 -5 %, level1 / stride-2 convs +6..+24 %: a net loss)
template <int PROD> struct GProd { static constexpr int warps = PROD == 0 ? 16 : 8; };
// warps: [producers | 4 epilogue warps | B loader, MMA issuer, 2 idle] -- every role group is a whole warpgroup, so
// fp32 mode can move registers between them with setmaxnreg (the band-drain epilogue keeps NT accumulators per thread)
template <int PROD> struct GThreads { static constexpr int value = (GProd<PROD>::warps + 4 + 4) * 32; };
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
constexpr int S_TMA_WARPS = 4;                           // bulk copies issued by ONE warp serialise (~530 clk
                                                         // each, tests/cuda/tma_bw_probe.cu): spread them over 4
// epilogue warps per TMEM lane quadrant (each takes half the columns).  Measured with 4 (16 epilogue warps of 96
// registers): 1x1 128 -> 64 97 -> 82 us, but 128 -> 128 3x3 131 -> 143 us and the head conv 708 -> 817 us: stays 2.
constexpr int S_EPI_NH = 2;
// warps: [4 TMA | MMA issuer, 3 idle | 4 * S_EPI_NH epilogue] -- whole warpgroups, so fp32 mode can hand the registers
// of the first two groups to the epilogue warps (setmaxnreg): at 128 registers the band-drain epilogue spilled
constexpr int S_W_MMA = S_TMA_WARPS, S_W_EPI = 8;
constexpr int S_THREADS = (S_W_EPI + 4 * S_EPI_NH) * 32;
// per-tile staging of the epilogue's scale / shift rows in shared memory: 2 buffers x [scale | shift] x 256 columns
constexpr int EPI_SS_BYTES = 2 * 2 * 256 * 4;
constexpr int S_MAX_SB = 24;                             // B ring slots (resident weights: one per K block)

enum { PROD_DCN = 0, PROD_STRIDE = 1, PROD_SMALLC = 2 };

struct ConvP {
  View x, res, y;
  const float* om;
  const unsigned char* wpack;
  const float* scale;
  const float* shift;
  float* yf;
  long long ldyf;
  int n_valid;
  int P;                 // output rows = B*(Ho+2)*(Wo+2)
  int Ho, Wo;
  int taps, KC, nkb;     // shift: taps in {1,9}; nkb = K blocks per tile
  int NT, n_tiles, m_tiles;
  int SA, SB, a_rows;
  int act, epi, tmem_cols;
  int acc_r;             // D0 accumulators per stage (k steps rotate over them: shorter fp32 chains)
  int acc_stages;        // TMEM accumulator stages (2 = epilogue overlaps the next tile)
  int nslots, nd1;       // fp32 mode: D0 band slots (2..6) and D1 buffers (1..2) in TMEM, NT columns each
  int wide;              // fp32 mode: band slots are [D0 | D1] (2 NT columns) filled by A_hi x [B_hi | B_lo] + A_lo x B_hi
  int slot_cols;         // TMEM columns per band slot: NT, or 2 NT when wide
  int stride;
  int spk;               // shift kernel: super-pixel weights, zero K steps of the neighbour taps skipped (EPI_SP2SC)
  int sx;                // gather kernels: column stride of the input anchor (== stride, except the super-pixel stem: 4)
  int in_Wp, in_Hp;      // input frame (gather kernels)
  int seg_groups;        // SMALLC: 16-byte groups per segment (8 or 4)
  int seg_off[16];       // SMALLC: input row offset of segment j of K block kb at [kb*2 + j]
  int vec8;              // SMALLC: rows are only 8-byte aligned (C = 4)
  unsigned long long magic_wp, magic_hp;   // ceil(2^64 / (Wo+2)), ceil(2^64 / (Ho+2)): exact m / d for m < 2^32
  int b_resident;        // gather kernels: all weight blocks stay in shared memory (n_tiles == 1)
  int stg_bytes;         // EPI_PL: bytes of the epilogue's store staging tile (one 64-channel chunk x planes) at the head of smem
  // EPI_HEADS (fp32 mode): the 1x1 output convolutions of the heads fused into the epilogue of the 64 -> 3 x 256 head
  // convolution (base_model.py:121-135, :190-199).  Tiles are walked M-major (one CTA takes all N tiles of an M tile),
  // head h = h_tiles consecutive N tiles; w2: per head [hid][w2_stride[h]] fp32 (row k = the output weights of hidden
  // channel k, zero padded), heads back to back at w2_off[h] floats; b2 [n_heads][8]; out2[h]: fp32 [B, nout[h], Ho, Wo]
  int m_major, n_heads, h_tiles, sig_mask, w2_floats, heads_bytes;
  const float* w2;
  const float* b2;
  float* out2[4];
  int nout[4], w2_off[4], w2_stride[4];
  int dbg;               // tools/conv_bench.py: 1 = skip A copies, 2 = skip B copies, 4 = skip epilogue math
  // dcn_tile_kernel: output tiles of 8 x 16 pixels; the input tile + halo is staged in shared memory
  int tile2d, tiles_x, tiles_y, halo, LW, LH, xt_plane;
};

// dcn_tile_kernel: output row (padded-frame index) of TMEM lane `row` of tile t
__device__ __forceinline__ int tile_row_m(const ConvP& p, int t, int row) {
  const int per = p.tiles_x * p.tiles_y;
  const int b = t / per, r = t - b * per;
  const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
  return (b * (p.Ho + 2) + ty * 8 + (row >> 4) + 1) * (p.Wo + 2) + tx * 16 + (row & 15) + 1;
}

static int g_dbg = 0;
int debug_flags() { return g_dbg; }

// m -> (px, py, b) of the zero-bordered output frame, without integer division
__device__ __forceinline__ void decode_row(const ConvP& p, int m, int& px, int& py, int& b) {
  const uint32_t t = (uint32_t)__umul64hi((unsigned long long)(uint32_t)m, p.magic_wp);
  px = m - (int)t * (p.Wo + 2);
  b = (int)__umul64hi((unsigned long long)t, p.magic_hp);
  py = (int)t - b * (p.Ho + 2);
}
// tile `it` of this CTA: N-major round-robin over all `total` tiles (the caller's count: dcn_tile_kernel walks 2-D
// tiles, not p.m_tiles), or (m_major) all N tiles of M tile blockIdx.x + q * gridDim.x
__device__ __forceinline__ bool tile_at(const ConvP& p, int total, int it, int& nt, int& m0, int& t) {
  if (p.m_major) {
    const int q = it / p.n_tiles;
    nt = it - q * p.n_tiles;
    const int mt = (int)blockIdx.x + q * (int)gridDim.x;
    m0 = mt * TM;
    t = mt * p.n_tiles + nt;
    return mt < p.m_tiles;
  }
  t = (int)blockIdx.x + it * (int)gridDim.x;
  nt = t % p.n_tiles;
  m0 = (t / p.n_tiles) * TM;
  return t < total;
}
static unsigned long long magic_u64(unsigned d) { return d <= 1 ? ~0ull : (~0ull / d) + 1; }

__host__ __device__ constexpr uint32_t idesc_f16_f32(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// one 64-wide K block: 4 K steps, NS*(NS+1)/2 MMAs each.  tacc = [D0_0 .. D0_{R-1}, D1], NT columns
// each.  The tensor core TRUNCATES when it adds a K=16 dot product into the fp32 accumulator
// (measured: -2.7e-5 relative on an all-positive K=4608 sum), so in the fp32-parity mode the hi*hi
// products rotate over R accumulators that the epilogue adds in round-to-nearest fp32.
template <int NS> struct AccR { static constexpr int value = NS == 2 ? 3 : 1; };

// rb = (index of this block's first K step) mod R
// KMASK: bit k set = K step k is issued (super-pixel weights: the K steps of a neighbour tap that are zero by
// construction are skipped; FIRST then requires bit 0, R == 1)
template <int NS, bool FIRST, int R = AccR<NS>::value, int KMASK = 15>
__device__ __forceinline__ void issue_kblock(uint32_t a_addr, uint32_t a_plane, uint32_t b_addr, uint32_t b_plane,
                                             uint32_t tacc, uint32_t NT, uint32_t idesc, int rb) {
  // everything but the 14-bit start-address field of a descriptor is constant: build the four
  // descriptors once, then only add 2 (= 32 bytes >> 4) per K step.  The issuing thread is the
  // serial bottleneck of small-N tiles (tests/cuda/mma_rate_probe.cu: every branch / loop trip
  // around the MMAs costs ~100 clk), so callers issue whole groups of K blocks straight-line.
  const uint64_t ah = smem_desc_sw128(a_addr), bh = smem_desc_sw128(b_addr);
  const uint64_t al = smem_desc_sw128(a_addr + a_plane), bl = smem_desc_sw128(b_addr + b_plane);
  const uint32_t tD1 = tacc + (uint32_t)R * NT;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!((KMASK >> k) & 1)) continue;
    int r = 0;
    if (R > 1) { r = rb + k; r = r >= R ? r - R : r; r = r >= R ? r - R : r; }
    const uint32_t tD0 = tacc + (uint32_t)r * NT;
    mma_bf16_ss(tD0, ah + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc, (FIRST && k < R) ? 0u : 1u);
    if (NS == 2) {
      mma_bf16_ss(tD1, ah + (uint64_t)(2 * k), bl + (uint64_t)(2 * k), idesc, (FIRST && k == 0) ? 0u : 1u);
      mma_bf16_ss(tD1, al + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc, 1u);
    }
  }
}

// one band of the shift kernel: TAPS K blocks that share an A band (descriptors one row apart),
// each with its own B slot; slots are released as soon as their MMAs retire
// SPK (3 taps, R == 1): super-pixel weights (planes.superpixel_weight, 4 pixels x 16 channels): the left neighbour
// contributes only its last pixel (K step 3), the right one only its first (K step 0) -- 6 K steps per band, not 12.
// The centre tap goes first so that the overwriting first MMA of a tile is a full one.
template <int NS, bool FIRST, int TAPS, int R, int SPK = 0>
__device__ __forceinline__ void issue_band(uint32_t a_base, uint32_t a_plane, const uint32_t (&b_addr)[3], uint32_t b_plane,
                                           uint32_t tacc, uint32_t NT, uint32_t idesc, int rb, uint64_t* const (&b_rel)[3],
                                           uint64_t* a_rel) {
  if constexpr (SPK && TAPS == 3) {
    issue_kblock<NS, FIRST, R, 15>(a_base + 128u, a_plane, b_addr[1], b_plane, tacc, NT, idesc, rb);
    mma_commit(b_rel[1]);
    issue_kblock<NS, false, R, 8>(a_base, a_plane, b_addr[0], b_plane, tacc, NT, idesc, rb);
    mma_commit(b_rel[0]);
    issue_kblock<NS, false, R, 1>(a_base + 256u, a_plane, b_addr[2], b_plane, tacc, NT, idesc, rb);
    mma_commit(b_rel[2]);
  } else {
#pragma unroll
    for (int dx = 0; dx < TAPS; ++dx) {
      if (dx == 0) issue_kblock<NS, FIRST, R>(a_base, a_plane, b_addr[0], b_plane, tacc, NT, idesc, rb);
      else issue_kblock<NS, false, R>(a_base + (uint32_t)dx * 128u, a_plane, b_addr[dx], b_plane, tacc, NT, idesc, rb + dx);
      mma_commit(b_rel[dx]);
    }
  }
  mma_commit(a_rel);
}

// ---- fp32 mode (NS = 2), band-drain accumulation.  The tensor core truncates every time it adds a K = 16 dot
// product into the fp32 accumulator (~1 ulp per add, biased towards zero), so a K = 4608 chain loses ~1e-5 -- ten
// times what an FFMA loop loses, and the synthetic network amplifies it.  Here a D0 (hi*hi) accumulator only ever
// sums ONE band of 12 K steps; the epilogue warps, idle during the main loop anyway, pull every finished band out
// of TMEM and add it to their registers in round-to-nearest fp32 while the MMA warp fills the other D0 slot.  D1
// (hi*lo + lo*hi, scaled 2^-11) runs over the whole tile: its truncation is 2^-11 times smaller.
// TMEM columns: [D0 slot 0 | D0 slot 1 | D1 of even tiles | D1 of odd tiles], NT each.
constexpr int BAND_KSTEPS = 12;
template <int KMASK = 15>
__device__ __forceinline__ void issue_kblock_f32(uint32_t a_addr, uint32_t a_plane, uint32_t b_addr, uint32_t b_plane,
                                                 uint32_t tD0, uint32_t tD1, uint32_t idesc, uint32_t acc0, uint32_t acc1) {
  const uint64_t ah = smem_desc_sw128(a_addr), bh = smem_desc_sw128(b_addr);
  const uint64_t al = smem_desc_sw128(a_addr + a_plane), bl = smem_desc_sw128(b_addr + b_plane);
  constexpr int K0 = (KMASK & 1) ? 0 : (KMASK & 2) ? 1 : (KMASK & 4) ? 2 : 3;       // first issued K step carries acc0 / acc1
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!((KMASK >> k) & 1)) continue;
    mma_bf16_ss(tD0, ah + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc, k == K0 ? acc0 : 1u);
    mma_bf16_ss(tD1, ah + (uint64_t)(2 * k), bl + (uint64_t)(2 * k), idesc, k == K0 ? acc1 : 1u);
    mma_bf16_ss(tD1, al + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc, 1u);
  }
}
template <int TAPS, int SPK = 0>
__device__ __forceinline__ void issue_band_f32(uint32_t a_base, uint32_t a_plane, const uint32_t (&b_addr)[3], uint32_t b_plane,
                                               uint32_t tD0, uint32_t tD1, uint32_t idesc, uint32_t acc0, uint32_t acc1,
                                               uint64_t* const (&b_rel)[3], uint64_t* a_rel) {
  if constexpr (SPK && TAPS == 3) {
    issue_kblock_f32<15>(a_base + 128u, a_plane, b_addr[1], b_plane, tD0, tD1, idesc, acc0, acc1);     // centre tap first
    mma_commit(b_rel[1]);
    issue_kblock_f32<8>(a_base, a_plane, b_addr[0], b_plane, tD0, tD1, idesc, 1u, 1u);
    mma_commit(b_rel[0]);
    issue_kblock_f32<1>(a_base + 256u, a_plane, b_addr[2], b_plane, tD0, tD1, idesc, 1u, 1u);
    mma_commit(b_rel[2]);
  } else {
#pragma unroll
    for (int dx = 0; dx < TAPS; ++dx) {
      issue_kblock_f32(a_base + (uint32_t)dx * 128u, a_plane, b_addr[dx], b_plane, tD0, tD1, idesc, dx == 0 ? acc0 : 1u,
                       dx == 0 ? acc1 : 1u);
      mma_commit(b_rel[dx]);
    }
  }
  mma_commit(a_rel);
}

// ---- fp32 mode, wide form: two MMAs per K step instead of three.  The hi and lo planes of a weight tile lie back to
// back in shared memory (NT rows of 128 B each, same swizzle phase), so ONE MMA of N = 2 NT columns computes
// A_hi x [B_hi | B_lo] into a band slot [D0 | D1] and a second one of N = NT adds A_lo x B_hi to the D1 half: A_hi is
// fetched from shared memory once instead of twice (operand bytes per K step: 24 -> 20 KB at NT = 128, 18 -> 14 KB at
// 64, 15 -> 11 KB at 32 -- the SS-mode operand fetch is what bounds these kernels), and the issuing thread has a
// third fewer instructions.  D1 is drained with its band (the epilogue adds d0 + d1 * 2^-11), so there is no
// tile-long D1 buffer and no D1 hand-shake.
template <int KMASK = 15>
__device__ __forceinline__ void issue_kblock_f32w(uint32_t a_addr, uint32_t a_plane, uint32_t b_addr, uint32_t tS, uint32_t NT,
                                                  uint32_t idesc_w, uint32_t idesc_n, uint32_t acc0) {
  const uint64_t ah = smem_desc_sw128(a_addr), bh = smem_desc_sw128(b_addr);
  const uint64_t al = smem_desc_sw128(a_addr + a_plane);
  constexpr int K0 = (KMASK & 1) ? 0 : (KMASK & 2) ? 1 : (KMASK & 4) ? 2 : 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!((KMASK >> k) & 1)) continue;
    mma_bf16_ss(tS, ah + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc_w, k == K0 ? acc0 : 1u);
    mma_bf16_ss(tS + NT, al + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc_n, 1u);
  }
}
template <int TAPS, int SPK = 0>
__device__ __forceinline__ void issue_band_f32w(uint32_t a_base, uint32_t a_plane, const uint32_t (&b_addr)[3], uint32_t tS,
                                                uint32_t NT, uint32_t idesc_w, uint32_t idesc_n, uint32_t acc0,
                                                uint64_t* const (&b_rel)[3], uint64_t* a_rel) {
  if constexpr (SPK && TAPS == 3) {
    issue_kblock_f32w<15>(a_base + 128u, a_plane, b_addr[1], tS, NT, idesc_w, idesc_n, acc0);     // centre tap first
    mma_commit(b_rel[1]);
    issue_kblock_f32w<8>(a_base, a_plane, b_addr[0], tS, NT, idesc_w, idesc_n, 1u);
    mma_commit(b_rel[0]);
    issue_kblock_f32w<1>(a_base + 256u, a_plane, b_addr[2], tS, NT, idesc_w, idesc_n, 1u);
    mma_commit(b_rel[2]);
  } else {
#pragma unroll
    for (int dx = 0; dx < TAPS; ++dx) {
      issue_kblock_f32w(a_base + (uint32_t)dx * 128u, a_plane, b_addr[dx], tS, NT, idesc_w, idesc_n, dx == 0 ? acc0 : 1u);
      mma_commit(b_rel[dx]);
    }
  }
  mma_commit(a_rel);
}

// explicit shared-memory accesses (pointers derived from the dynamic smem base otherwise compile to
// generic LD / ST)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// shared -> global bulk copy (TMA engine), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int NH>
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 2, %0;" ::"n"(NH * 128) : "memory"); }

// ------------------------------------------------------------------------------ epilogue
// 16 accumulator columns of this thread's row: sum of the R hi*hi accumulators (+ D1 * 2^-11)
template <int NS>
__device__ __forceinline__ void read_acc16(uint32_t tacc, int NT, int R, int c0, float (&o)[16]) {
  uint32_t d[16];
  tmem_ld16(tacc + (uint32_t)c0, d);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(d[j]);
  for (int r = 1; r < R; ++r) {
    tmem_ld16(tacc + (uint32_t)(r * NT + c0), d);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] += __uint_as_float(d[j]);
  }
  if (NS == 2) {
    tmem_ld16(tacc + (uint32_t)(R * NT + c0), d);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = fmaf(__uint_as_float(d[j]), LO_INV, o[j]);
  }
}

// One thread = one output row (TMEM lane).  tacc: TMEM address of D0 incl. the lane base.
// stage_s != 0 (EPI_PL, tile fully inside the view): rows are written to a shared-memory image of the
// output tile (same swizzle as global) and ONE thread bulk-stores each 16 KB chunk-plane; border rows
// are stored as zeros, which is what they hold anyway.
// The epilogue is instruction-latency bound (one warp per SM sub-partition walks ~300 dependent
// instructions per 16 columns), so the shift kernel runs NH = 2 warps per TMEM lane quadrant: warp
// `half` takes its share of the 16-column groups of every 64-column chunk.
struct EpiRow {
  int m, px, py, b;
  bool inP, valid, has_res, staged;
  uint32_t srow, sxor;
};

// accumulator columns [c0, c0+16) of this thread's row: sum of the R hi*hi accumulators (+ D1 * 2^-11);
// all TMEM reads are issued back to back
template <int NS>
__device__ __forceinline__ void epi_read(const ConvP& p, uint32_t tacc, int c0, float (&o)[16]) {
  const int NT = p.NT, R = p.acc_r;
  uint32_t d0[16], d1[16], d2[16], dl[16];
  tmem_ld16(tacc + (uint32_t)c0, d0);
  if (R == 3) {
    tmem_ld16(tacc + (uint32_t)(NT + c0), d1);
    tmem_ld16(tacc + (uint32_t)(2 * NT + c0), d2);
  }
  if (NS == 2) tmem_ld16(tacc + (uint32_t)(R * NT + c0), dl);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(d0[j]);
  if (R == 3) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { o[j] += __uint_as_float(d1[j]); o[j] += __uint_as_float(d2[j]); }
  }
  if (NS == 2) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = fmaf(__uint_as_float(dl[j]), LO_INV, o[j]);
  }
}

// EPI_SP2SC: the GEMM rows are super-pixels (4 pixels x 16 channels) of an Ho x Wo frame; pixel j of row e lands in
// the SC output frame of Ho x 4*Wo pixels
__device__ __forceinline__ long long sp2sc_row(const ConvP& p, const EpiRow& e, int j) {
  return ((long long)e.b * (p.Ho + 2) + e.py) * (4 * p.Wo + 2) + 4 * (e.px - 1) + 1 + j;
}

// dual-stem epilogue of one output pixel: out[c] = relu(bn_a(a[c])) + relu(bn_b(b[c])), c < 16 (dla.py:325-331)
__device__ __forceinline__ void stem_pixel(const ConvP& p, const float (&fa)[16], const float (&fb)[16], float (&lo8)[8],
                                           float (&hi8)[8]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float va = fmaf(fa[j], __ldg(p.scale + j), __ldg(p.shift + j));
    const float vb = fmaf(fb[j], __ldg(p.scale + 16 + j), __ldg(p.shift + 16 + j));
    const float o = fmaxf(va, 0.f) + fmaxf(vb, 0.f);
    if (j < 8) lo8[j] = o; else hi8[j - 8] = o;
  }
}

// the same with the 32 scale / shift values read from their shared-memory staging (fp32_epilogue_loop)
__device__ __forceinline__ void stem_pixel_ss(uint32_t ss, const float (&fa)[16], const float (&fb)[16], float (&o16)[16]) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const uint4 sa = lds128(ss + 16u * j4), sb = lds128(ss + 64u + 16u * j4);
    const uint4 ha = lds128(ss + 1024u + 16u * j4), hb = lds128(ss + 1088u + 16u * j4);
    const uint32_t s_a[4] = {sa.x, sa.y, sa.z, sa.w}, s_b[4] = {sb.x, sb.y, sb.z, sb.w};
    const uint32_t h_a[4] = {ha.x, ha.y, ha.z, ha.w}, h_b[4] = {hb.x, hb.y, hb.z, hb.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = 4 * j4 + i;
      const float va = fmaf(fa[j], __uint_as_float(s_a[i]), __uint_as_float(h_a[i]));
      const float vb = fmaf(fb[j], __uint_as_float(s_b[i]), __uint_as_float(h_b[i]));
      o16[j] = fmaxf(va, 0.f) + fmaxf(vb, 0.f);
    }
  }
}

// scale / shift / residual / activation / store of 16 columns held in registers
// SS: scale[16] at shared-memory address ss, shift[16] at ss + 1024 (staged by epi_stage_ss)
template <int NS, bool SS = false>
__device__ __forceinline__ void epi_finish(const ConvP& p, const EpiRow& e, int n, float (&o)[16], uint32_t ss = 0u) {
  uint4 rv[2][NS];
  if (e.has_res) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ch = n + 8 * h;
#pragma unroll
      for (int pl = 0; pl < NS; ++pl)
        rv[h][pl] = __ldg(reinterpret_cast<const uint4*>(p.res.base + pl_offset(p.res, pl, ch >> 6, e.m, (ch & 63) >> 3)));
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 a, b4;
    if constexpr (SS) {
      const uint4 ua = lds128(ss + 16u * j), ub = lds128(ss + 1024u + 16u * j);
      a = make_float4(__uint_as_float(ua.x), __uint_as_float(ua.y), __uint_as_float(ua.z), __uint_as_float(ua.w));
      b4 = make_float4(__uint_as_float(ub.x), __uint_as_float(ub.y), __uint_as_float(ub.z), __uint_as_float(ub.w));
    } else {
      a = __ldg(reinterpret_cast<const float4*>(p.scale + n) + j);
      b4 = __ldg(reinterpret_cast<const float4*>(p.shift + n) + j);
    }
    o[4 * j] = fmaf(o[4 * j], a.x, b4.x); o[4 * j + 1] = fmaf(o[4 * j + 1], a.y, b4.y);
    o[4 * j + 2] = fmaf(o[4 * j + 2], a.z, b4.z); o[4 * j + 3] = fmaf(o[4 * j + 3], a.w, b4.w);
  }
  if (p.epi == SGTA_EPI_PL || p.epi == SGTA_EPI_SC || p.epi == SGTA_EPI_SP2SC) {
    if (!e.staged && !e.valid) return;
    if (e.has_res) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float f[8];
        decode8<NS>(rv[h][0], rv[h][NS - 1], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[8 * h + j] += f[j];
      }
    }
    if (p.act == SGTA_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    if (!e.staged) {       // 16 channels = one aligned 32-byte sector per plane in every layout: STG.256
      if (p.epi == SGTA_EPI_PL) pl_store16<NS>(p.y, n >> 6, e.m, (n & 63) >> 3, o);
      else if (p.epi == SGTA_EPI_SC) sc_store16<NS>(p.y, e.m, n, o);
      else sc_store16<NS>(p.y, sp2sc_row(p, e, n >> 4), 0, o);
      return;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {          // staged: the tile's shared-memory image, bulk-stored by the caller
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = o[8 * h + j];
      const int ch = n + 8 * h;
      uint4 e0, e1;
      encode8<NS>(f, e0, e1);
      if (!e.valid) { e0 = make_uint4(0, 0, 0, 0); e1 = e0; }
      const uint32_t a = e.srow + ((((uint32_t)(ch & 63) >> 3) ^ e.sxor) << 4);
      sts128(a, e0);
      if (NS == 2) sts128(a + 16384u, e1);
    }
  } else if (p.epi == SGTA_EPI_F32ROWS) {
    if (!e.inP) return;
    float* dst = p.yf + (size_t)e.m * p.ldyf + n;
    if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
        stg256(dst + 8 * j, make_uint4(__float_as_uint(o[8 * j]), __float_as_uint(o[8 * j + 1]), __float_as_uint(o[8 * j + 2]), __float_as_uint(o[8 * j + 3])),
               make_uint4(__float_as_uint(o[8 * j + 4]), __float_as_uint(o[8 * j + 5]), __float_as_uint(o[8 * j + 6]), __float_as_uint(o[8 * j + 7])));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(dst)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
  } else {   // SGTA_EPI_NCHW: fp32 [B, n_valid, Ho, Wo]
    if (!e.valid) return;
    const size_t hw = (size_t)p.Ho * p.Wo;
    float* dst = p.yf + ((size_t)e.b * p.n_valid + n) * hw + (size_t)(e.py - 1) * p.Wo + (e.px - 1);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (n + j < p.n_valid) {
        float v = o[j];
        if (p.act == SGTA_ACT_RELU) v = fmaxf(v, 0.f);
        else if (p.act == SGTA_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
        dst[(size_t)j * hw] = v;
      }
    }
  }
}

template <int NS>
__device__ __forceinline__ void epilogue_group(const ConvP& p, const EpiRow& e, uint32_t tacc, int c0, int n, uint32_t ss) {
  float o[16];
  epi_read<NS>(p, tacc, c0, o);
  epi_finish<NS, true>(p, e, n, o, ss + (uint32_t)c0 * 4u);
}

// ---- fp32 band-drain epilogue (see issue_kblock_f32): one thread = one output row (TMEM lane) and NG 16-column
// groups of it, starting at group g0 (NH epilogue warps per lane quadrant split the columns between them)
struct Ring {
  uint32_t i, ph;
  int n;
  __device__ __forceinline__ void next() { if ((int)++i == n) { i = 0; ph ^= 1u; } }
};
__device__ __forceinline__ void epi_row_setup(const ConvP& p, int m, EpiRow& e) {
  e.m = m;
  decode_row(p, e.m, e.px, e.py, e.b);
  e.inP = e.m < p.P;
  e.valid = e.inP && e.px >= 1 && e.px <= p.Wo && e.py >= 1 && e.py <= p.Ho;
  const bool pl_like = p.epi == SGTA_EPI_PL || p.epi == SGTA_EPI_SC || p.epi == SGTA_EPI_SP2SC;
  e.has_res = pl_like && e.valid && p.res.base != nullptr;
  e.staged = false;
  e.srow = 0; e.sxor = 0;
}
// ---- EPI_HEADS: hidden = relu(scale * acc + shift) stays in registers; out[o] += hidden[k] * w2[k][o] over the thread's
// columns of the N tile (fp32 FMA), accumulated over the h_tiles N tiles of a head (the CTA walks them back to back:
// m_major); after the last one the NH column halves of a row meet through shared memory (two addends: order-free),
// bias + optional sigmoid, fp32 NCHW store (lane = pixel: coalesced per output channel).  The 768-channel hidden map
// never reaches HBM (it was 856 MB written + 945 MB read back per step of 32 frame pairs) and the three 1x1 launches go.
template <int NG, int NH>
__device__ __forceinline__ void heads_tile(const ConvP& p, int nt, int m0, int q, int half, int lane, int g0, uint32_t ss,
                                           float (&acc)[NG][16], float (&h2)[8], float* w2s) {
  const int hd = nt / p.h_tiles, sub = nt - hd * p.h_tiles;
  const int stride = p.w2_stride[hd];
  const uint32_t wbase = smem_u32(w2s + p.w2_off[hd]);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if ((g0 + g) * 16 >= p.NT) continue;
    const uint32_t ssg = ss + (uint32_t)((g0 + g) * 64);
    const int k0 = sub * p.NT + (g0 + g) * 16;                 // hidden channel of column 0 of this group, inside the head
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const uint4 ua = lds128(ssg + 16u * j4), ub = lds128(ssg + 1024u + 16u * j4);
      const uint32_t sc4[4] = {ua.x, ua.y, ua.z, ua.w}, sh4[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = 4 * j4 + i;
        const float v = fmaxf(fmaf(acc[g][j], __uint_as_float(sc4[i]), __uint_as_float(sh4[i])), 0.f);
        const uint32_t wa = wbase + (uint32_t)((k0 + j) * stride) * 4u;
        if (stride == 8) {
          const uint4 w0 = lds128(wa), w1 = lds128(wa + 16u);
          h2[0] = fmaf(v, __uint_as_float(w0.x), h2[0]); h2[1] = fmaf(v, __uint_as_float(w0.y), h2[1]);
          h2[2] = fmaf(v, __uint_as_float(w0.z), h2[2]); h2[3] = fmaf(v, __uint_as_float(w0.w), h2[3]);
          h2[4] = fmaf(v, __uint_as_float(w1.x), h2[4]); h2[5] = fmaf(v, __uint_as_float(w1.y), h2[5]);
          h2[6] = fmaf(v, __uint_as_float(w1.z), h2[6]); h2[7] = fmaf(v, __uint_as_float(w1.w), h2[7]);
        } else if (stride == 4) {
          const uint4 w0 = lds128(wa);
          h2[0] = fmaf(v, __uint_as_float(w0.x), h2[0]); h2[1] = fmaf(v, __uint_as_float(w0.y), h2[1]);
          h2[2] = fmaf(v, __uint_as_float(w0.z), h2[2]); h2[3] = fmaf(v, __uint_as_float(w0.w), h2[3]);
        } else {
          uint32_t wx, wy;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wx), "=r"(wy) : "r"(wa));
          h2[0] = fmaf(v, __uint_as_float(wx), h2[0]); h2[1] = fmaf(v, __uint_as_float(wy), h2[1]);
        }
      }
    }
  }
  if (sub != p.h_tiles - 1) return;
  // last N tile of the head: the column halves of each row meet
  const int row = q * 32 + lane;
  float* xch = w2s + p.w2_floats + row * 8;
  if (NH > 1 && half != 0) {
    sts128(smem_u32(xch), make_uint4(__float_as_uint(h2[0]), __float_as_uint(h2[1]), __float_as_uint(h2[2]), __float_as_uint(h2[3])));
    sts128(smem_u32(xch) + 16u, make_uint4(__float_as_uint(h2[4]), __float_as_uint(h2[5]), __float_as_uint(h2[6]), __float_as_uint(h2[7])));
  }
  if (NH > 1) asm volatile("bar.sync 3, %0;" ::"n"(NH * 128) : "memory");
  if (half == 0) {
    if (NH > 1) {
      const uint4 x0 = lds128(smem_u32(xch)), x1 = lds128(smem_u32(xch) + 16u);
      h2[0] += __uint_as_float(x0.x); h2[1] += __uint_as_float(x0.y); h2[2] += __uint_as_float(x0.z); h2[3] += __uint_as_float(x0.w);
      h2[4] += __uint_as_float(x1.x); h2[5] += __uint_as_float(x1.y); h2[6] += __uint_as_float(x1.z); h2[7] += __uint_as_float(x1.w);
    }
    EpiRow e;
    epi_row_setup(p, m0 + row, e);
    if (e.valid) {
      const int no = p.nout[hd];
      const size_t hw = (size_t)p.Ho * p.Wo;
      float* dst = p.out2[hd] + (size_t)e.b * no * hw + (size_t)(e.py - 1) * p.Wo + (e.px - 1);
      const bool sig = (p.sig_mask >> hd) & 1;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        if (o < no) {
          float v = h2[o] + __ldg(p.b2 + hd * 8 + o);
          if (sig) v = 1.f / (1.f + expf(-v));
          dst[(size_t)o * hw] = v;
        }
      }
    }
  }
  if (NH > 1) asm volatile("bar.sync 3, %0;" ::"n"(NH * 128) : "memory");      // the exchange buffer is free again
#pragma unroll
  for (int o = 0; o < 8; ++o) h2[o] = 0.f;
}

// scale / shift of the N tile staged in shared memory by the epilogue warps themselves (two buffers, refreshed when the
// N tile changes): read per 16-column group from global memory at the end of a tile they were L1 misses under the TMA
// traffic -- half of the epilogue warps' stall samples on the 64 -> 768 head convolution.  et = index of the calling
// thread among the NH * 128 epilogue threads, which all call this once per tile.  Returns the buffer's address.
template <int NH>
__device__ __forceinline__ uint32_t epi_stage_ss(const ConvP& p, int nt, int& ss_nt, uint32_t& ss_buf, float* epi_ss, int et) {
  if (nt != ss_nt) {
    ss_nt = nt;
    ss_buf ^= 1u;
    const bool is_stem = p.epi == SGTA_EPI_STEM || p.epi == SGTA_EPI_STEM_SP;
    const int ss_n = is_stem ? 32 : p.NT;               // the dual stem has 32 scale / shift values whatever its N
    for (int c = et; c < ss_n; c += NH * 128) {
      epi_ss[ss_buf * 512 + c] = __ldg(p.scale + nt * p.NT + c);
      epi_ss[ss_buf * 512 + 256 + c] = __ldg(p.shift + nt * p.NT + c);
    }
    asm volatile("bar.sync 3, %0;" ::"n"(NH * 128) : "memory");
  }
  return smem_u32(epi_ss) + ss_buf * 2048u;
}
template <int NG, int NH, bool BACKOFF>
__device__ __forceinline__ void fp32_epilogue_loop(const ConvP& p, uint32_t tmem, int total, int nbands, int q, int half,
                                                   int lane, uint64_t* slot_full, uint64_t* slot_empty, uint64_t* d1_empty,
                                                   float* epi_ss) {
  const int NT = p.NT, g0 = half * NG;
  const bool wide = p.wide != 0;
  const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
  Ring rs{0u, 0u, p.nslots}, rd{0u, 0u, p.nd1};
  int ss_nt = -1;
  uint32_t ss_buf = 0;
  // EPI_HEADS: the fused 1x1 weights sit behind the scale / shift staging, then the half-to-half exchange buffer
  float h2[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) h2[o] = 0.f;
  float* w2s = epi_ss + EPI_SS_BYTES / 4;
  if (p.epi == SGTA_EPI_HEADS) {
    for (int i = (half * 4 + q) * 32 + lane; i < p.w2_floats; i += NH * 128) w2s[i] = __ldg(p.w2 + i);
    asm volatile("bar.sync 3, %0;" ::"n"(NH * 128) : "memory");
  }
  for (int it = 0, nt, m0, t; tile_at(p, total, it, nt, m0, t); ++it) {
    const uint32_t ss = epi_stage_ss<NH>(p, nt, ss_nt, ss_buf, epi_ss, (half * 4 + q) * 32 + lane);
    float acc[NG][16];
    float acc1[NG <= 2 ? NG : 1][16];                 // wide form, NG <= 2: the D1 halves of the band slots
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int j = 0; j < 16; ++j) { acc[g][j] = 0.f; if (NG <= 2) acc1[NG <= 2 ? g : 0][j] = 0.f; }
    for (int bnd = 0; bnd < nbands; ++bnd) {
      if (BACKOFF) mbar_wait_backoff(&slot_full[rs.i], rs.ph);
      else mbar_wait(&slot_full[rs.i], rs.ph);
      tc_fence_after();
      if (!(p.dbg & 4)) {
        const uint32_t ts = lane_base + rs.i * (uint32_t)p.slot_cols;
        if (wide) {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if ((g0 + g) * 16 < NT) {
              if constexpr (NG <= 2) {                     // D1 summed on its own, scaled in once per tile
                uint32_t d[16], e1[16];
                tmem_ld16(ts + (uint32_t)((g0 + g) * 16), d);
                tmem_ld16(ts + (uint32_t)(NT + (g0 + g) * 16), e1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) { acc[g][j] += __uint_as_float(d[j]); acc1[g][j] += __uint_as_float(e1[j]); }
              } else if constexpr (NG <= 4) {
                uint32_t d[16], e1[16];
                tmem_ld16(ts + (uint32_t)((g0 + g) * 16), d);
                tmem_ld16(ts + (uint32_t)(NT + (g0 + g) * 16), e1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[g][j] += fmaf(__uint_as_float(e1[j]), LO_INV, __uint_as_float(d[j]));
              } else {                                     // 128 accumulators per thread: one 16-register buffer
                uint32_t d[16];
                tmem_ld16(ts + (uint32_t)((g0 + g) * 16), d);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[g][j] += __uint_as_float(d[j]);
                tmem_ld16(ts + (uint32_t)(NT + (g0 + g) * 16), d);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[g][j] = fmaf(__uint_as_float(d[j]), LO_INV, acc[g][j]);
              }
            }
          }
        } else {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if ((g0 + g) * 16 < NT) {
              uint32_t d[16];
              tmem_ld16(ts + (uint32_t)((g0 + g) * 16), d);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[g][j] += __uint_as_float(d[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&slot_empty[rs.i]);
      rs.next();
    }
    if (wide) {
      if constexpr (NG <= 2) {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[g][j] = fmaf(acc1[g][j], LO_INV, acc[g][j]);
      }
    } else {
      if (!(p.dbg & 4)) {
        const uint32_t td = lane_base + (uint32_t)(p.nslots + (int)rd.i) * (uint32_t)NT;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if ((g0 + g) * 16 < NT) {
            uint32_t d[16];
            tmem_ld16(td + (uint32_t)((g0 + g) * 16), d);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[g][j] = fmaf(__uint_as_float(d[j]), LO_INV, acc[g][j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d1_empty[rd.i]);
      rd.next();
    }
    if (p.dbg & (4 | 8192)) continue;                       // 8192: drain TMEM but skip the math and the stores
    if (p.epi == SGTA_EPI_HEADS) {
      if constexpr (NG <= 4) heads_tile<NG, NH>(p, nt, m0, q, half, lane, g0, ss, acc, h2, w2s);
      continue;
    }
    EpiRow e;
    epi_row_setup(p, p.tile2d ? tile_row_m(p, t / p.n_tiles, q * 32 + lane) : m0 + q * 32 + lane, e);
    if (p.epi == SGTA_EPI_STEM || p.epi == SGTA_EPI_STEM_SP) {
      // N = 32 per output pixel: [16 image-conv | 16 heat-map-conv] columns; NH = 1 here.  EPI_STEM: one pixel per row
      // (N = 32) -> SC row; EPI_STEM_SP: the row is a super-pixel of 4 pixels (N = 128) -> 32-byte group j of the PL row
      if (NG >= 2 && e.valid) {
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) {
          if (j * 32 < NT) {
            float o16[16];
            stem_pixel_ss(ss, acc[NG >= 2 ? 2 * j : 0], acc[NG >= 2 ? 2 * j + 1 : 0], o16);
            if (p.epi == SGTA_EPI_STEM) sc_store16<2>(p.y, e.m, 0, o16);
            else pl_store16<2>(p.y, 0, e.m, 2 * j, o16);
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int g = 0; g < NG; ++g)
      if ((g0 + g) * 16 < NT) epi_finish<2, true>(p, e, nt * NT + (g0 + g) * 16, acc[g], ss + (uint32_t)((g0 + g) * 64));
  }
}
template <int NH, bool BACKOFF, int MAXNG = 8 / NH>
__device__ __forceinline__ void fp32_epilogue(const ConvP& p, uint32_t tmem, int total, int nbands, int q, int half, int lane,
                                              uint64_t* slot_full, uint64_t* slot_empty, uint64_t* d1_empty, float* epi_ss) {
  const int per = ((p.NT + 15) / 16 + NH - 1) / NH;          // 16-column groups per thread (the launcher keeps it <= MAXNG)
  if (per <= 1) fp32_epilogue_loop<1, NH, BACKOFF>(p, tmem, total, nbands, q, half, lane, slot_full, slot_empty, d1_empty, epi_ss);
  else if (per <= 2) fp32_epilogue_loop<2, NH, BACKOFF>(p, tmem, total, nbands, q, half, lane, slot_full, slot_empty, d1_empty, epi_ss);
  else if (per <= 4 || MAXNG <= 4) fp32_epilogue_loop<4, NH, BACKOFF>(p, tmem, total, nbands, q, half, lane, slot_full, slot_empty, d1_empty, epi_ss);
  else if constexpr (MAXNG > 4) fp32_epilogue_loop<8, NH, BACKOFF>(p, tmem, total, nbands, q, half, lane, slot_full, slot_empty, d1_empty, epi_ss);
}

// NH epilogue warps per lane quadrant; `half` in [0, NH) is this warp's share
template <int NS, int NH>
__device__ __forceinline__ void epilogue_tile(const ConvP& p, uint32_t tacc, int m0, int n0, int row, uint32_t stage_s,
                                              int half, uint64_t* release, uint32_t ss, int m_abs = -1) {
  const int NT = p.NT;
  EpiRow e;
  e.m = m_abs >= 0 ? m_abs : m0 + row;
  decode_row(p, e.m, e.px, e.py, e.b);
  e.inP = e.m < p.P;
  e.valid = e.inP && e.px >= 1 && e.px <= p.Wo && e.py >= 1 && e.py <= p.Ho;
  if (p.epi == SGTA_EPI_STEM || p.epi == SGTA_EPI_STEM_SP) {
    // 32 columns per output pixel: [16 image-conv | 16 heat-map-conv]   (dla.py:325-331); see fp32_epilogue_loop
    if (half != 0) return;
    const int R = p.acc_r;
    for (int j = 0; j * 32 < NT; ++j) {
      float fa[16], fb[16];
      __syncwarp();
      read_acc16<NS>(tacc, NT, R, 32 * j, fa);
      read_acc16<NS>(tacc, NT, R, 32 * j + 16, fb);
      if (e.valid) {
        float lo8[8], hi8[8];
        stem_pixel(p, fa, fb, lo8, hi8);
        if (p.epi == SGTA_EPI_STEM) {
          sc_store8<NS>(p.y, e.m, 0, lo8);
          sc_store8<NS>(p.y, e.m, 8, hi8);
        } else {
          pl_store8<NS>(p.y, 0, e.m, 2 * j, lo8);
          pl_store8<NS>(p.y, 0, e.m, 2 * j + 1, hi8);
        }
      }
    }
    return;
  }
  const bool pl_like = p.epi == SGTA_EPI_PL || p.epi == SGTA_EPI_SC || p.epi == SGTA_EPI_SP2SC;
  e.has_res = pl_like && e.valid && p.res.base != nullptr;
  e.staged = stage_s != 0 && p.epi == SGTA_EPI_PL && m0 + TM <= p.P;
  e.srow = stage_s + (uint32_t)row * 128u;
  e.sxor = (uint32_t)((p.y.guard + e.m) & 7);
  const int G = (NT < 64 ? NT : 64) >> 4;              // 16-column groups per chunk: 1, 2 or 4
  const bool leader = row == 0 && half == 0;
  for (int c0 = 0; c0 < NT; c0 += 16) {
    __syncwarp();                      // tcgen05.ld is warp-collective: reconverge after the stores
    const int n = n0 + c0;
    if (e.staged && (c0 & 63) == 0) {  // staging tile free again: the previous chunk's bulk store has read it
      if (leader) bulk_wait_read0();
      epi_bar<NH>();
    }
    const int gi = (c0 >> 4) & (G - 1);
    const bool mine = NH == 1 || (G == 1 ? half == 0 : (gi * NH) / G == half);
    if (mine) epilogue_group<NS>(p, e, tacc, c0, n, ss);
    if (e.staged && (c0 & 63) == 48) {
      fence_async_smem();
      epi_bar<NH>();
      if (leader) {
#pragma unroll
        for (int pl = 0; pl < NS; ++pl)
          bulk_s2g(p.y.base + ((((size_t)pl * p.y.nchunks + p.y.chunk0 + (n >> 6)) * p.y.rows + p.y.guard + m0) << 7),
                   stage_s + (uint32_t)pl * 16384u, 16384u);
        bulk_commit();
      }
    }
  }
}

__device__ __forceinline__ unsigned char* align1024(unsigned char* p) {
  return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}
__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* slot, int cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// =============================================================================== shift kernel
template <int NS, int R, int SPK = 0>
__global__ void __launch_bounds__(S_THREADS, 1) conv_shift_kernel(const __grid_constant__ ConvP p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align1024(smem_raw);
  const int NT = p.NT;
  const uint32_t a_plane = (uint32_t)p.a_rows * 128u, a_stage = a_plane * NS;
  const uint32_t b_plane = (uint32_t)NT * 128u, b_stage = b_plane * NS;
  unsigned char* sA = smem + p.stg_bytes;
  unsigned char* sB = sA + (size_t)p.SA * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.SB * b_stage);
  uint64_t *a_full = bars, *a_empty = bars + 8, *b_full = bars + 56, *b_empty = bars + 56 + S_MAX_SB;
  uint64_t *acc_full = bars + 32, *acc_empty = bars + 34;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);
  uint64_t *slot_full = bars + 38, *slot_empty = bars + 44;     // fp32 mode: up to 6 D0 band slots (acc_empty = D1 buffers)
  float* epi_ss = reinterpret_cast<float*>(bars + 128);         // scale / shift staging of the fp32 epilogue

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(&a_full[s], NS); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < S_MAX_SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4 * S_EPI_NH); }
    for (int s = 0; s < 6; ++s) { mbar_init(&slot_full[s], 1); mbar_init(&slot_empty[s], 4 * S_EPI_NH); }
    fence_mbar_init();
  }
  if (warp == S_W_MMA) tmem_alloc_dyn(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                          // barriers / TMEM are set up: now the inputs must be complete
  const uint32_t tmem = *tmem_slot;
  const uint32_t acc_stride = (uint32_t)((p.acc_r + NS - 1) * NT);
  const int total = p.m_tiles * p.n_tiles;
  const int Wp = p.x.W + 2;

  const int nb = p.taps == 9 ? 3 : 1;
  const int ntap_b = p.taps == 9 ? 3 : 1;

  if (warp < S_TMA_WARPS) {
    if constexpr (NS == 2) reg_dec<40>();
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producers
      // every copy (one plane of an A band, one B tile) belongs to a FIXED warp per ring slot, so the
      // rounds of a slot are issued in order by one thread (a parity wait two rounds early would pass)
      // ring positions are tracked incrementally (a runtime % or / per K block costs ~100 clk on
      // a single thread, more than the MMAs it feeds)
      int st = 0, sb = 0;
      uint32_t aph = 1, bph = 1;           // parity to wait for on the *_empty barriers
      const bool skipA = p.dbg & 1, skipB = p.dbg & 2;
      for (int it = 0, nt, m0, t; tile_at(p, total, it, nt, m0, t); ++it) {
        const unsigned char* wt = p.wpack + (size_t)nt * p.nkb * b_stage;
        for (int kc = 0; kc < p.KC; ++kc) {
          for (int band = 0; band < nb; ++band) {
            const long long s = (long long)p.x.guard + m0 + (p.taps == 9 ? (band - 1) * Wp - 1 : 0);
            const long long s8 = s & ~7ll;
#pragma unroll
            for (int pl = 0; pl < NS; ++pl) {
              if (((st * NS + pl) & (S_TMA_WARPS - 1)) != warp) continue;
              mbar_wait(&a_empty[st], aph);
              if (skipA) { mbar_arrive(&a_full[st]); continue; }
              mbar_arrive_expect_tx(&a_full[st], a_plane);
              bulk_g2s(sA + (size_t)st * a_stage + pl * a_plane,
                       p.x.base + ((((size_t)pl * p.x.nchunks + p.x.chunk0 + kc) * p.x.rows + s8) << 7), a_plane,
                       &a_full[st]);
            }
            if (++st == p.SA) { st = 0; aph ^= 1u; }
            for (int dx = 0; dx < ntap_b; ++dx) {
              if (((sb + 2) & (S_TMA_WARPS - 1)) == warp) {
                mbar_wait(&b_empty[sb], bph);
                // resident weights (SB == nkb): every slot is filled during this CTA's first tile and only re-armed later
                if (skipB || (p.b_resident && it != 0)) mbar_arrive(&b_full[sb]);
                else {
                  const int tap = band * ntap_b + dx;
                  mbar_arrive_expect_tx(&b_full[sb], b_stage);
                  bulk_g2s(sB + (size_t)sb * b_stage, wt + ((size_t)tap * p.KC + kc) * b_stage, b_stage, &b_full[sb]);
                }
              }
              if (++sb == p.SB) { sb = 0; bph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == S_W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (whole warp
    // walks the loop and waits; one elected lane issues, so operands stay in uniform registers)
    if constexpr (NS == 2) reg_dec<40>();
    const uint32_t idesc = NS == 2 ? idesc_f16_f32(TM, NT) : idesc_bf16_f32(TM, NT);
    const uint32_t idesc_w = idesc_f16_f32(TM, 2 * NT);
    int st = 0, sb = 0;
    uint32_t aph = 0, bph = 0;             // parity to wait for on the *_full barriers
    uint32_t as = 0, accph = 1;            // accumulator stage and parity of its acc_empty barrier
    Ring rs{0u, 1u, p.nslots}, rd{0u, 1u, p.nd1};   // fp32 mode: D0 band slots / D1 buffers (parity of their *_empty)
    const uint32_t a_smem = smem_u32(sA), b_smem = smem_u32(sB);
    for (int it = 0, nt_, m0, t_; tile_at(p, total, it, nt_, m0, t_); ++it) {
      uint32_t tacc = 0, tD0 = 0, tD1 = 0, sl = 0;
      if constexpr (NS == 2) {
        if (!p.wide) {
          mbar_wait(&acc_empty[rd.i], rd.ph);                        // the epilogue has read this D1 buffer
          tD1 = tmem + (uint32_t)(p.nslots + (int)rd.i) * (uint32_t)NT;
          rd.next();
        }
      } else {
        mbar_wait(&acc_empty[as], accph);
        tacc = tmem + as * acc_stride;
      }
      int rb = 0;                          // 1x1: K step rotation start (3x3 bands advance by 12 = 0 mod 3)
      int ks = 0;                          // fp32 mode: K steps already in the open D0 slot
      bool first = true;
      for (int kc = 0; kc < p.KC; ++kc) {
        for (int band = 0; band < nb; ++band) {
          const long long s = (long long)p.x.guard + m0 + (p.taps == 9 ? (band - 1) * Wp - 1 : 0);
          const uint32_t off = (uint32_t)(s & 7ll);
          mbar_wait(&a_full[st], aph);
          const uint32_t a_base = a_smem + (uint32_t)st * a_stage + off * 128u;
          uint32_t b_addr[3];
          uint64_t* b_rel[3];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            if (dx < ntap_b) {
              mbar_wait(&b_full[sb], bph);
              b_addr[dx] = b_smem + (uint32_t)sb * b_stage;
              b_rel[dx] = &b_empty[sb];
              if (++sb == p.SB) { sb = 0; bph ^= 1u; }
            } else {
              b_addr[dx] = b_smem;
              b_rel[dx] = &b_empty[0];
            }
          }
          if constexpr (NS == 2) {
            if (ks == 0) {
              sl = rs.i;
              mbar_wait(&slot_empty[sl], rs.ph);                      // the epilogue has drained this D0 slot
              tD0 = tmem + sl * (uint32_t)p.slot_cols;
              rs.next();
            }
            tc_fence_after();
            if (elect_one()) {
              const uint32_t acc0 = ks ? 1u : 0u, acc1 = first ? 0u : 1u;
              if (p.wide) {
                if (ntap_b == 3) issue_band_f32w<3, SPK>(a_base, a_plane, b_addr, tD0, (uint32_t)NT, idesc_w, idesc, acc0, b_rel, &a_empty[st]);
                else issue_band_f32w<1>(a_base, a_plane, b_addr, tD0, (uint32_t)NT, idesc_w, idesc, acc0, b_rel, &a_empty[st]);
              } else {
                if (ntap_b == 3) issue_band_f32<3, SPK>(a_base, a_plane, b_addr, b_plane, tD0, tD1, idesc, acc0, acc1, b_rel, &a_empty[st]);
                else issue_band_f32<1>(a_base, a_plane, b_addr, b_plane, tD0, tD1, idesc, acc0, acc1, b_rel, &a_empty[st]);
              }
            }
            __syncwarp();
            ks += SPK ? 6 : ntap_b * 4;
            const bool last = kc == p.KC - 1 && band == nb - 1;
            if (ks >= BAND_KSTEPS || last) {
              if (elect_one()) mma_commit(&slot_full[sl]);
              __syncwarp();
              ks = 0;
            }
          } else {
            tc_fence_after();
            if (elect_one()) {
              if (ntap_b == 3) {
                if (first) issue_band<NS, true, 3, R, SPK>(a_base, a_plane, b_addr, b_plane, tacc, (uint32_t)NT, idesc, 0, b_rel, &a_empty[st]);
                else issue_band<NS, false, 3, R, SPK>(a_base, a_plane, b_addr, b_plane, tacc, (uint32_t)NT, idesc, 0, b_rel, &a_empty[st]);
              } else {
                if (first) issue_band<NS, true, 1, R>(a_base, a_plane, b_addr, b_plane, tacc, (uint32_t)NT, idesc, 0, b_rel, &a_empty[st]);
                else issue_band<NS, false, 1, R>(a_base, a_plane, b_addr, b_plane, tacc, (uint32_t)NT, idesc, rb, b_rel, &a_empty[st]);
              }
            }
            __syncwarp();
          }
          first = false;
          if (R == 3 && ntap_b == 1) rb = rb == 2 ? 0 : rb + 1;
          if (++st == p.SA) { st = 0; aph ^= 1u; }
        }
      }
      if constexpr (NS != 2) {
        if (elect_one()) mma_commit(&acc_full[as]);
        __syncwarp();
        if (p.acc_stages == 2) { as ^= 1u; if (as == 0) accph ^= 1u; } else accph ^= 1u;
      }
    }
  } else if (warp >= S_W_EPI) {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3, half = (warp - S_W_EPI) >> 2;
    if constexpr (NS == 2) {
      reg_inc<208>();
      // fp32 mode: drain every finished band into registers, then D1, then math + stores (under the next tile's MMAs)
      const int steps = p.KC * nb;                              // issue_band calls per tile, ntap_b * 4 K steps each
      const int per = BAND_KSTEPS / (SPK ? 6 : ntap_b * 4);     // ... per D0 band: 1 (3x3), 3 (1x1) or 2 (super-pixel 3x3)
      fp32_epilogue<S_EPI_NH, false>(p, tmem, total, (steps + per - 1) / per, q, half, lane, slot_full, slot_empty, acc_empty, epi_ss);
    } else {
      uint32_t as = 0, accph = 0, ss_buf = 0;
      int ss_nt = -1;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int nt = t % p.n_tiles, m0 = (t / p.n_tiles) * TM;
        const uint32_t ss = epi_stage_ss<S_EPI_NH>(p, nt, ss_nt, ss_buf, epi_ss, (half * 4 + q) * 32 + lane);
        mbar_wait_backoff(&acc_full[as], accph);
        tc_fence_after();
        if (!(p.dbg & 4))
          epilogue_tile<NS, S_EPI_NH>(p, tmem + as * acc_stride + ((uint32_t)(q * 32) << 16), m0, nt * NT, q * 32 + lane,
                                      p.stg_bytes ? smem_u32(smem) : 0u, half, nullptr, ss);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        if (p.acc_stages == 2) { as ^= 1u; if (as == 0) accph ^= 1u; } else accph ^= 1u;
      }
      if (q == 0 && half == 0 && lane == 0) bulk_wait_all0();
    }
  }
  else {
    if constexpr (NS == 2) reg_dec<40>();             // idle warps of the MMA warpgroup
  }
  tc_fence_before();
  __syncthreads();
  if (warp == S_W_MMA) tmem_dealloc_dyn(tmem, p.tmem_cols);
}

// =============================================================================== gather kernel
// WIDE (fp32 mode only): N tiles of more than 64 columns -- the epilogue threads then hold up to 128 accumulators.
// Register budget at launch: 768 x 80 (DCN) or 512 x 128 (plain gathers); the B loader / MMA warpgroup gives up all
// but 40 per thread, which takes the DCN epilogue warpgroup to 120 (152 when the producers go down to 72 as well)
// and the plain gathers' to 208.
template <int PROD, int NS, int WIDE = 0>
__global__ void __launch_bounds__(GThreads<PROD>::value, 1) conv_gather_kernel(const __grid_constant__ ConvP p) {
  constexpr int G_PROD_WARPS = GProd<PROD>::warps;
  constexpr int W_EPI = G_PROD_WARPS, W_BLD = G_PROD_WARPS + 4, W_MMA = G_PROD_WARPS + 5;
  constexpr int RPL = TM / (G_PROD_WARPS * 4);          // rows per lane and K block: 4 (8 warps) or 2 (16 warps)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align1024(smem_raw);
  const int NT = p.NT;
  constexpr uint32_t a_plane = TM * 128u, a_bytes = a_plane * NS;
  const uint32_t b_plane = (uint32_t)NT * 128u, b_bytes = b_plane * NS;
  // ring stage = [A | B]; with resident weights the ring holds A only and all B blocks sit behind it
  const uint32_t stage_bytes = p.b_resident ? a_bytes : a_bytes + b_bytes;
  unsigned char* ring = smem + p.stg_bytes;            // [epilogue store staging | operand ring | ...]
  unsigned char* sBres = ring + (size_t)p.SA * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBres + (p.b_resident ? (size_t)p.nkb * b_bytes : 0));
  uint64_t *full = bars, *empty = bars + 8, *acc_full = bars + 16, *acc_empty = bars + 18;
  uint64_t* b_ready = bars + 21;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint64_t *slot_full = bars + 22, *slot_empty = bars + 28;              // fp32 mode: up to 6 D0 band slots (acc_empty = D1 buffers)
  unsigned char* tab = reinterpret_cast<unsigned char*>(bars + 48);      // DCN sampling table (9*128*32 B)
  float* epi_ss = reinterpret_cast<float*>(tab + (PROD == PROD_DCN ? 9 * TM * 32 : 0));     // fp32 epilogue: scale / shift staging

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(&full[s], G_PROD_WARPS + (p.b_resident ? 0 : 1)); mbar_init(&empty[s], 1); }
    mbar_init(b_ready, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    for (int s = 0; s < 6; ++s) { mbar_init(&slot_full[s], 1); mbar_init(&slot_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc_dyn(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const uint32_t acc_stride = (uint32_t)((p.acc_r + NS - 1) * NT);
  // tile2d (DCN, fp32 mode, n_tiles == 1): M tiles are blocks of 8 x 16 pixels (tile_row_m) instead of 128 consecutive rows
  const int total = p.tile2d ? p.tiles_x * p.tiles_y * p.x.B : p.m_tiles * p.n_tiles;
  const int nkb = p.nkb;

  if (warp < G_PROD_WARPS) {
    // ================================================================== A producers
    if constexpr (NS == 2 && WIDE && PROD == PROD_DCN) reg_dec<72>();
    // lane = (sub, c8): 8 lanes cover one 128-byte row, a warp-wide load covers 4 whole rows
    const int sub = lane >> 3, c8 = lane & 7;
    const int iWp = p.in_Wp, iHp = p.in_Hp;
    int pst = 0;           // ring slot and parity of its empty barrier
    uint32_t pph = 1;
    auto wait_stage = [&]() -> unsigned char* {
      mbar_wait(&empty[pst], pph);
      return ring + (size_t)pst * stage_bytes;
    };
    auto publish = [&]() {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[pst]);
      if (++pst == p.SA) { pst = 0; pph ^= 1u; }
    };
    const size_t plane_bytes = ((size_t)p.x.nchunks * p.x.rows) << 7;     // PL inputs
    uint32_t soff[RPL];    // swizzled byte offset of this thread's 16-byte group in each of its rows
    int rr[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      rr[i] = warp * (4 * RPL) + i * 4 + sub;
      soff[i] = sw128_offset((uint32_t)rr[i], (uint32_t)c8);
    }

    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int m0 = (t / p.n_tiles) * TM;
      if constexpr (PROD == PROD_DCN) {
        // ---- per tile: sampling table in shared memory, one entry per (tap, row): the byte offsets
        // of the four corner rows inside a chunk (16-byte group 0, swizzle folded in: a lane only
        // XORs its own group index) and the four mask*bilinear weights; every lane of a row reads
        // them back (broadcast).  The producers are instruction-bound (tools/dcn_bench.py: the same
        // time with the global loads removed), so everything row-invariant lives here.
        // A warp only ever reads the entries of ITS OWN 4 * RPL rows, so it builds exactly those (9 taps x 4 * RPL rows)
        // and nothing but __syncwarp orders the table: the 512 producers no longer meet at two CTA-wide barriers per
        // tile (ncu source page: 15 % of the kernel's stall samples sat on them) and drift as far as the ring allows.
        uint4* tab_o = reinterpret_cast<uint4*>(tab);
        float4* tab_w = reinterpret_cast<float4*>(tab + 9 * TM * 16);
        __syncwarp();
        for (int e2 = lane; e2 < 9 * 4 * RPL; e2 += 32) {
          const int tap = e2 / (4 * RPL), row = warp * (4 * RPL) + (e2 - tap * (4 * RPL));
          const int e = tap * TM + row;
          const int m = p.tile2d ? tile_row_m(p, t, row) : m0 + row;
          int px, py, b;
          decode_row(p, m, px, py, b);
          const bool ok = m < p.P && px >= 1 && px <= p.Wo && py >= 1 && py <= p.Ho;
          float dy = 0.f, dx = 0.f, ml = 0.f;
          if (ok) {
            const float* om = p.om + (size_t)m * 32;
            dy = __ldg(om + 2 * tap); dx = __ldg(om + 2 * tap + 1); ml = __ldg(om + 18 + tap);
          }
          // padded-frame coordinates (unpadded + 1): the zero border implements "zero outside"
          const int ky = tap / 3, kx = tap - 3 * ky;
          float sy = (float)(py - 1 + ky) + dy;
          float sx = (float)(px - 1 + kx) + dx;
          const bool in = ok && sy > 0.f && sy < (float)(p.Ho + 1) && sx > 0.f && sx < (float)(p.Wo + 1);
          const float msk = in ? 1.f / (1.f + __expf(-ml)) : 0.f;
          sy = in ? sy : 0.f;
          sx = in ? sx : 0.f;
          const float yf = floorf(sy), xf = floorf(sx);
          const float ly = sy - yf, lx = sx - xf, hy = 1.f - ly, hx = 1.f - lx;
          const uint32_t r0 = (uint32_t)(min(b, p.x.B - 1) * iHp * iWp + (int)yf * iWp + (int)xf + p.x.guard);
          const uint32_t r1 = r0 + 1u, r2 = r0 + (uint32_t)iWp, r3 = r2 + 1u;
          tab_o[e] = make_uint4((r0 << 7) | ((r0 & 7u) << 4), (r1 << 7) | ((r1 & 7u) << 4),
                                (r2 << 7) | ((r2 & 7u) << 4), (r3 << 7) | ((r3 & 7u) << 4));
          tab_w[e] = make_float4(msk * hy * hx, msk * hy * lx, msk * ly * hx, msk * ly * lx);
        }
        __syncwarp();
        const uint32_t gx = (uint32_t)c8 << 4;
        const uint32_t tab_s = smem_u32(tab);
        for (int kc = 0; kc < p.KC; ++kc) {
          const unsigned char* xk = p.x.base + (((size_t)(p.x.chunk0 + kc) * p.x.rows) << 7);
          const unsigned char* xl = xk + plane_bytes;
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t sA_s = smem_u32(wait_stage());
            if (p.dbg & 1) { publish(); continue; }
            constexpr int GRW = 2;                      // rows blended together (loads in flight: GRW x 4 corners x NS)
#pragma unroll
            for (int ih = 0; ih < RPL / GRW; ++ih) {
              uint4 v[GRW][4][NS];
              float4 w[GRW];
#pragma unroll
              for (int ii = 0; ii < GRW; ++ii) {
                const int i = ih * GRW + ii;
                const uint4 o4 = lds128(tab_s + (uint32_t)(tap * TM + rr[i]) * 16u);
                const uint4 w4 = lds128(tab_s + (uint32_t)(9 * TM + tap * TM + rr[i]) * 16u);
                w[ii] = make_float4(__uint_as_float(w4.x), __uint_as_float(w4.y), __uint_as_float(w4.z), __uint_as_float(w4.w));
                const uint32_t off[4] = {o4.x ^ gx, o4.y ^ gx, o4.z ^ gx, o4.w ^ gx};
#pragma unroll
                for (int cn = 0; cn < 4; ++cn) {
                  v[ii][cn][0] = __ldg(reinterpret_cast<const uint4*>(xk + off[cn]));
                  if (NS == 2) v[ii][cn][NS - 1] = __ldg(reinterpret_cast<const uint4*>(xl + off[cn]));
                }
              }
#pragma unroll
              for (int ii = 0; ii < GRW; ++ii) {
                const int i = ih * GRW + ii;
                const float wc[4] = {w[ii].x, w[ii].y, w[ii].z, w[ii].w};
                uint4 e0, e1 = make_uint4(0, 0, 0, 0);
                if (NS == 2) {
                  // hi plane blended in fp32; lo plane (a 2^-11-scaled correction) blended in packed
                  // fp16 (its rounding lands at 2^-21 relative); the convex combination of fp16-range
                  // values needs no clamp before the re-split
                  float H[8];
                  __half2 L[4];
#pragma unroll
                  for (int cn = 0; cn < 4; ++cn) {
                    const uint32_t hv[4] = {v[ii][cn][0].x, v[ii][cn][0].y, v[ii][cn][0].z, v[ii][cn][0].w};
                    const uint32_t lv[4] = {v[ii][cn][1].x, v[ii][cn][1].y, v[ii][cn][1].z, v[ii][cn][1].w};
                    const __half2 wh = __float2half2_rn(wc[cn]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float2 hf = unpack_h2(hv[j]);
                      const __half2 lh = *reinterpret_cast<const __half2*>(&lv[j]);
                      if (cn == 0) {
                        H[2 * j] = wc[0] * hf.x; H[2 * j + 1] = wc[0] * hf.y;
                        L[j] = __hmul2(wh, lh);
                      } else {
                        H[2 * j] = fmaf(wc[cn], hf.x, H[2 * j]); H[2 * j + 1] = fmaf(wc[cn], hf.y, H[2 * j + 1]);
                        L[j] = __hfma2(wh, lh, L[j]);
                      }
                    }
                  }
                  uint32_t eh[4], el[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 lf = __half22float2(L[j]);
                    const float a = fmaf(lf.x, LO_INV, H[2 * j]), b2 = fmaf(lf.y, LO_INV, H[2 * j + 1]);
                    const __half2 h = __floats2half2_rn(a, b2);
                    const float2 hb = __half22float2(h);
                    eh[j] = *reinterpret_cast<const uint32_t*>(&h);
                    el[j] = pack_h2((a - hb.x) * LO_SCALE, (b2 - hb.y) * LO_SCALE);
                  }
                  e0 = make_uint4(eh[0], eh[1], eh[2], eh[3]);
                  e1 = make_uint4(el[0], el[1], el[2], el[3]);
                } else {
                  float o[8];
#pragma unroll
                  for (int cn = 0; cn < 4; ++cn) {
                    const uint32_t bv[4] = {v[ii][cn][0].x, v[ii][cn][0].y, v[ii][cn][0].z, v[ii][cn][0].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float2 f = unpack_b2(bv[j]);
                      if (cn == 0) { o[2 * j] = wc[0] * f.x; o[2 * j + 1] = wc[0] * f.y; }
                      else { o[2 * j] = fmaf(wc[cn], f.x, o[2 * j]); o[2 * j + 1] = fmaf(wc[cn], f.y, o[2 * j + 1]); }
                    }
                  }
                  e0 = make_uint4(pack_b2(o[0], o[1]), pack_b2(o[2], o[3]), pack_b2(o[4], o[5]), pack_b2(o[6], o[7]));
                }
                sts128(sA_s + soff[i], e0);
                if (NS == 2) sts128(sA_s + a_plane + soff[i], e1);
              }
            }
            publish();
          }
        }
      }
    }
    if constexpr (PROD != PROD_DCN) {
      // ---- plain gathers (STRIDE: PL input, SMALLC: SC input).  The K blocks of all tiles of this
      // CTA form one flat stream; loads run PD blocks ahead of the shared-memory stores so PD
      // blocks of global-load latency overlap per warp (8 warps per SM cannot hide it otherwise).
      constexpr int PD = 3;
      uint4 v[PD][RPL][NS];
      const int C2 = p.x.nchunks * 2;                   // SC: bytes per input row
      const size_t sc_plane = (size_t)p.x.rows * C2;
      const int seg = c8 / p.seg_groups, within = c8 - seg * p.seg_groups;
      const int my_tiles = blockIdx.x < total ? (total - 1 - blockIdx.x) / gridDim.x + 1 : 0;
      const long long nblk = (long long)my_tiles * nkb;
      int lt = blockIdx.x, lkb = 0;                     // load cursor: tile, K block inside the tile
      int anchor[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) anchor[i] = 0;
      auto set_tile = [&](int t) {
        const int m0 = (t / p.n_tiles) * TM;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          int px, py, b;
          decode_row(p, m0 + rr[i], px, py, b);
          b = min(b, p.x.B - 1);                          // border / tail rows: any in-bounds address
          const int oy = min(max(py - 1, 0), p.Ho - 1), ox = min(max(px - 1, 0), p.Wo - 1);
          anchor[i] = p.x.guard + b * iHp * iWp + oy * p.stride * iWp + ox * p.sx;
        }
      };
      auto load_block = [&](uint4 (&dst)[RPL][NS]) {
        if (lkb == 0) set_tile(lt);
        if (PROD == PROD_STRIDE) {
          // 3x3, pad 1, stride s: padded input coords of tap (ky,kx) = (oy*s+ky, ox*s+kx); kc outer
          const int kc = lkb / 9, tap = lkb - kc * 9;
          const int toff = (tap / 3) * iWp + tap % 3;
          const unsigned char* xk = p.x.base + (((size_t)(p.x.chunk0 + kc) * p.x.rows) << 7);
#pragma unroll
          for (int i = 0; i < RPL; ++i) {
            const int r = anchor[i] + toff;
            const unsigned char* src = xk + ((size_t)r << 7) + (((c8 ^ r) & 7) << 4);
#pragma unroll
            for (int pl = 0; pl < NS; ++pl) dst[i][pl] = __ldg(reinterpret_cast<const uint4*>(src + pl * plane_bytes));
          }
        } else {
          if (p.vec8) {
            // C = 4 (rows only 8-byte aligned): group c8 = [pixel c8 of segment 0 | pixel c8 of segment 1],
            // so the 8 lanes of a row read two contiguous 64-byte runs (the weight matrix uses the same order)
            const int so0 = p.seg_off[lkb * 2], so1 = p.seg_off[lkb * 2 + 1];
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
              const unsigned char* s0 = p.x.base + (size_t)(anchor[i] + so0) * C2 + c8 * 8;
              const unsigned char* s1 = p.x.base + (size_t)(anchor[i] + so1) * C2 + c8 * 8;
#pragma unroll
              for (int pl = 0; pl < NS; ++pl) {
                const uint2 a = __ldg(reinterpret_cast<const uint2*>(s0 + pl * sc_plane));
                const uint2 b2 = __ldg(reinterpret_cast<const uint2*>(s1 + pl * sc_plane));
                dst[i][pl] = make_uint4(a.x, a.y, b2.x, b2.y);
              }
            }
          } else {
            // group c8 of K block kb = 16 bytes at (anchor + seg_off[kb][seg]) * C*2 + within*16
            const int so = p.seg_off[lkb * 2 + seg];
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
              const unsigned char* src0 = p.x.base + (size_t)(anchor[i] + so) * C2 + within * 16;
#pragma unroll
              for (int pl = 0; pl < NS; ++pl) dst[i][pl] = __ldg(reinterpret_cast<const uint4*>(src0 + pl * sc_plane));
            }
          }
        }
        if (++lkb == nkb) { lkb = 0; lt += gridDim.x; }
      };
      auto store_block = [&](const uint4 (&src)[RPL][NS]) {
        const uint32_t sA = smem_u32(wait_stage());           // explicit STS (a generic pointer compiles to ST.E)
#pragma unroll
        for (int i = 0; i < RPL; ++i)
#pragma unroll
          for (int pl = 0; pl < NS; ++pl) sts128(sA + pl * a_plane + soff[i], src[i][pl]);
        publish();
      };
      if (p.dbg & 1) {
        for (long long j = 0; j < nblk; ++j) { wait_stage(); publish(); }
      } else {
        for (long long j0 = 0; j0 < nblk + PD - 1; j0 += PD) {
#pragma unroll
          for (int d = 0; d < PD; ++d) {
            const long long j = j0 + d;
            // slot d holds block j; it was stored (as block j - PD) PD-1 iterations ago
            if (j >= PD - 1 && j - (PD - 1) < nblk) store_block(v[(d + 1) % PD]);
            if (j < nblk) load_block(v[d]);
          }
        }
      }
    }
  } else if (warp >= W_BLD) {
   if constexpr (NS == 2) reg_dec<40>();
   if (warp == W_BLD) {
    // ==================================================================== B loader
    if (lane == 0) {
      if (p.b_resident) {
        // weights are loaded once per CTA; packed block order (tap, kc) is kept
        mbar_arrive_expect_tx(b_ready, (uint32_t)nkb * b_bytes);
        for (int kb = 0; kb < nkb; ++kb)
          bulk_g2s(sBres + (size_t)kb * b_bytes, p.wpack + (size_t)kb * b_bytes, b_bytes, b_ready);
      } else {
        int s = 0;
        uint32_t ph = 1;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
          const unsigned char* wt = p.wpack + (size_t)(t % p.n_tiles) * nkb * b_bytes;
          // producers walk (kc outer, tap inner) for DCN / STRIDE; packed blocks are (tap, kc)
          int kc = 0, tap = 0;
          for (int kb = 0; kb < nkb; ++kb) {
            const int blk = PROD == PROD_SMALLC ? kb : tap * p.KC + kc;
            mbar_wait(&empty[s], ph);
            if (p.dbg & 2) mbar_arrive(&full[s]);
            else {
              mbar_arrive_expect_tx(&full[s], b_bytes);
              bulk_g2s(ring + (size_t)s * stage_bytes + a_bytes, wt + (size_t)blk * b_bytes, b_bytes, &full[s]);
            }
            if (++s == p.SA) { s = 0; ph ^= 1u; }
            if (++tap == 9) { tap = 0; ++kc; }
          }
        }
      }
    }
   } else if (warp == W_MMA) {
    // ==================================================================== MMA issuer
    const uint32_t idesc = NS == 2 ? idesc_f16_f32(TM, NT) : idesc_bf16_f32(TM, NT);
    const uint32_t idesc_w = idesc_f16_f32(TM, 2 * NT);
    constexpr int R = AccR<NS>::value;
    constexpr int BAND_KB = BAND_KSTEPS / 4;                 // fp32 mode: K blocks per D0 band
    int s = 0;
    uint32_t ph = 0, as = 0, accph = 1;
    Ring rs{0u, 1u, p.nslots}, rd{0u, 1u, p.nd1};
    const uint32_t smem0 = smem_u32(ring), bres0 = smem_u32(sBres);
    if (p.b_resident) mbar_wait(b_ready, 0);
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      uint32_t tacc = 0, tD0 = 0, tD1 = 0, sl = 0;
      if constexpr (NS == 2) {
        if (!p.wide) {
          mbar_wait(&acc_empty[rd.i], rd.ph);
          tD1 = tmem + (uint32_t)(p.nslots + (int)rd.i) * (uint32_t)NT;
          rd.next();
        }
      } else {
        mbar_wait(&acc_empty[as], accph);
        tacc = tmem + as * acc_stride;
      }
      int rb = 0, kc = 0, tap = 0, kin = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[s], ph);
        const uint32_t a0 = smem0 + (uint32_t)s * stage_bytes;
        const int blk = PROD == PROD_SMALLC ? kb : tap * p.KC + kc;
        const uint32_t b0 = p.b_resident ? bres0 + (uint32_t)blk * b_bytes : a0 + a_bytes;
        if (++tap == 9) { tap = 0; ++kc; }
        if constexpr (NS == 2) {
          if (kin == 0) {
            sl = rs.i;
            mbar_wait(&slot_empty[sl], rs.ph);
            tD0 = tmem + sl * (uint32_t)p.slot_cols;
            rs.next();
          }
          tc_fence_after();
          if (elect_one()) {
            if (!(p.dbg & 16)) {
              if (p.wide) issue_kblock_f32w(a0, a_plane, b0, tD0, (uint32_t)NT, idesc_w, idesc, kin ? 1u : 0u);
              else issue_kblock_f32(a0, a_plane, b0, b_plane, tD0, tD1, idesc, kin ? 1u : 0u, kb ? 1u : 0u);
            }
            mma_commit(&empty[s]);
            if (kin + 1 == BAND_KB || kb == nkb - 1) mma_commit(&slot_full[sl]);
          }
          __syncwarp();
          if (++kin == BAND_KB || kb == nkb - 1) kin = 0;
        } else {
          tc_fence_after();
          if (elect_one()) {
            if (!(p.dbg & 16)) {
              if (kb == 0) issue_kblock<NS, true>(a0, a_plane, b0, b_plane, tacc, (uint32_t)NT, idesc, 0);
              else issue_kblock<NS, false>(a0, a_plane, b0, b_plane, tacc, (uint32_t)NT, idesc, rb);
            }
            mma_commit(&empty[s]);
          }
          __syncwarp();
        }
        if (R == 3) rb = rb == 2 ? 0 : rb + 1;
        if (++s == p.SA) { s = 0; ph ^= 1u; }
      }
      if constexpr (NS != 2) {
        if (elect_one()) mma_commit(&acc_full[as]);
        __syncwarp();
        if (p.acc_stages == 2) { as ^= 1u; if (as == 0) accph ^= 1u; } else accph ^= 1u;
      }
    }
   }
  } else {
    // ==================================================================== epilogue warps
    const int q = warp & 3;
    if constexpr (NS == 2) {
      if constexpr (PROD == PROD_DCN) { if constexpr (WIDE) reg_inc<152>(); else reg_inc<120>(); }
      else reg_inc<208>();
      constexpr int BAND_KB = BAND_KSTEPS / 4;
      fp32_epilogue<1, true, (WIDE || PROD != PROD_DCN) ? 8 : 4>(p, tmem, total, (nkb + BAND_KB - 1) / BAND_KB, q, 0, lane,
                                                                  slot_full, slot_empty, acc_empty, epi_ss);
    } else {
      uint32_t as = 0, accph = 0, ss_buf = 0;
      int ss_nt = -1;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int nt = t % p.n_tiles, m0 = (t / p.n_tiles) * TM;
        const uint32_t ss = epi_stage_ss<1>(p, nt, ss_nt, ss_buf, epi_ss, q * 32 + lane);
        mbar_wait_backoff(&acc_full[as], accph);
        tc_fence_after();
        if (!(p.dbg & 4))
          epilogue_tile<NS, 1>(p, tmem + as * acc_stride + ((uint32_t)(q * 32) << 16), m0, nt * NT, q * 32 + lane,
                               p.stg_bytes ? smem_u32(smem) : 0u, 0, nullptr, ss);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        if (p.acc_stages == 2) { as ^= 1u; if (as == 0) accph ^= 1u; } else accph ^= 1u;
      }
      if (q == 0 && lane == 0) bulk_wait_all0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc_dyn(tmem, p.tmem_cols);
}

// =============================================================================== DCN tile kernel
// The DeformConv layers on maps of >= 8 x 16 pixels (13 of the 16 at 384 x 384: dla.py:545).  Same GEMM, same
// B loader / MMA issuer / epilogue protocol as conv_gather_kernel<PROD_DCN>; what changes is where the producers
// gather from.  An M tile is a block of 8 x 16 output pixels, and for every 64-channel chunk the input block
// plus a halo of `halo` pixels ((8 + 2h) x (16 + 2h) rows of 128 B per plane) is staged in shared memory by bulk
// copies (one per image row and plane, issued by the 16 producer warps' lane 0).  The 4 x NS corner loads of a
// sample are LDS.128 from that tile: 128 B/clk instead of the ~64 B/clk an LDG.128 that touches four cache lines
// gets from L1, a fixed 29-cycle latency, no tag look-ups, and the unified L1 no longer has to stay large -- the
// x tile, the operand ring and the sampling table fill the 227 KB.  Samples whose corners leave the staged
// window (|offset| >= halo - 1) take the global-memory path of the old kernel, row by row.
constexpr int D_PROD_WARPS = 16;
constexpr int D_THREADS = 24 * 32;      // warps 0-15 producers | 16-19 epilogue | 20 B loader, 21 MMA issuer, 22-23 idle

// WIDE = 1: N tiles of 128 columns in fp32 mode -- the epilogue threads then hold 128 accumulators each and need
// 152 registers.  setmaxnreg moves registers only inside the CTA's own allocation (768 x 80 at launch): the B
// loader / MMA warpgroup gives up 40 per thread, which takes the epilogue warpgroup to 120 (enough for 64 columns);
// for 128 columns the producers go down to 72 as well.
template <int NS, int WIDE>
__global__ void __launch_bounds__(D_THREADS, 1) dcn_tile_kernel(const __grid_constant__ ConvP p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = align1024(smem_raw);
  const int NT = p.NT;
  constexpr uint32_t a_plane = TM * 128u, a_bytes = a_plane * NS;
  const uint32_t b_plane = (uint32_t)NT * 128u, b_bytes = b_plane * NS;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t xt_plane = (uint32_t)p.xt_plane;
  unsigned char* xt = smem;                                    // [plane][LH][LW][128 B], rows as they lie in HBM
  unsigned char* ring = smem + (size_t)xt_plane * NS;          // SA stages of [A | B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.SA * stage_bytes);
  uint64_t *full = bars, *empty = bars + 8, *acc_full = bars + 16, *acc_empty = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint64_t* xbar = bars + 21;
  uint64_t *slot_full = bars + 22, *slot_empty = bars + 28;
  unsigned char* tab = reinterpret_cast<unsigned char*>(bars + 48);      // sampling table (9*128*32 B)
  float* epi_ss = reinterpret_cast<float*>(tab + 9 * TM * 32);           // fp32 epilogue: scale / shift staging

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(&full[s], D_PROD_WARPS + 1); mbar_init(&empty[s], 1); }
    mbar_init(xbar, D_PROD_WARPS);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    for (int s = 0; s < 6; ++s) { mbar_init(&slot_full[s], 1); mbar_init(&slot_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == 21) tmem_alloc_dyn(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const int total = p.tiles_x * p.tiles_y * p.x.B;             // n_tiles == 1
  const int nkb = p.nkb;
  const int Wp = p.Wo + 2, Hp = p.Ho + 2;
  const int per_img = p.tiles_x * p.tiles_y;

  if (warp < D_PROD_WARPS) {
    // ================================================================== A producers
    if (WIDE) reg_dec<72>();
    const int sub = lane >> 3, c8 = lane & 7;
    int pst = 0;
    uint32_t pph = 1, xph = 0;
    const size_t plane_bytes = ((size_t)p.x.nchunks * p.x.rows) << 7;
    constexpr int RPL = TM / (D_PROD_WARPS * 4);               // 2 rows per lane and K block
    uint32_t soff[RPL];
    int rr[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      rr[i] = warp * (4 * RPL) + i * 4 + sub;
      soff[i] = sw128_offset((uint32_t)rr[i], (uint32_t)c8);
    }
    const uint32_t gx = (uint32_t)c8 << 4;
    const uint32_t tab_s = smem_u32(tab), xt_s = smem_u32(xt);
    const int h = p.halo, LW = p.LW, LH = p.LH;
    const uint32_t seg_bytes = (uint32_t)LW * 128u;
    const int ncopies = LH * NS;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int b = t / per_img, r_ = t - b * per_img;
      const int tyi = r_ / p.tiles_x, txi = r_ - tyi * p.tiles_x;
      const int y0 = tyi * 8, x0 = txi * 16;                   // un-padded origin of the output block
      // padded-frame row of local (0, 0) of the staged window, as a global row index of the view
      const long long g00 = (long long)p.x.guard + ((long long)b * Hp + (y0 + 1 - h)) * Wp + (x0 + 1 - h);
      for (int kc = 0; kc < p.KC; ++kc) {
        // every producer is done with the previous chunk's x tile (and, at kc == 0, with the previous table)
        asm volatile("bar.sync 1, %0;" ::"n"(D_PROD_WARPS * 32) : "memory");
        if (lane == 0) {
          int mine = 0;
          for (int i = warp; i < ncopies; i += D_PROD_WARPS) ++mine;
          if (mine) mbar_arrive_expect_tx(xbar, (uint32_t)mine * seg_bytes);
          else mbar_arrive(xbar);
          for (int i = warp; i < ncopies; i += D_PROD_WARPS) {
            const int pl = i / LH, ly = i - pl * LH;
            const unsigned char* src = p.x.base +
                ((((size_t)pl * p.x.nchunks + p.x.chunk0 + kc) * p.x.rows + (size_t)(g00 + (long long)ly * Wp)) << 7);
            bulk_g2s(xt + (size_t)pl * xt_plane + (size_t)ly * seg_bytes, src, seg_bytes, xbar);
          }
        }
        if (kc == 0) {
          // ---- sampling table of the tile: per (tap, row) the shared-memory addresses of the four corner rows
          // (plane 0, 16-byte group 0, swizzle of the row's HBM position folded in) and mask * bilinear weights;
          // bit 31 of .x flags a sample outside the staged window: .x then holds the global row of corner 0
          uint4* tab_o = reinterpret_cast<uint4*>(tab);
          float4* tab_w = reinterpret_cast<float4*>(tab + 9 * TM * 16);
          // (a warp builds the entries of its own 4 * RPL rows -- the only ones it reads: no CTA barrier after the build)
          for (int e2 = lane; e2 < 9 * 4 * RPL; e2 += 32) {
            const int tap = e2 / (4 * RPL), row = warp * (4 * RPL) + (e2 - tap * (4 * RPL));
            const int e = tap * TM + row;
            const int py = y0 + (row >> 4) + 1, px = x0 + (row & 15) + 1;       // padded coordinates
            const float* om = p.om + ((size_t)((long long)b * Hp + py) * Wp + px) * 32;
            const float dy = __ldg(om + 2 * tap), dx = __ldg(om + 2 * tap + 1), ml = __ldg(om + 18 + tap);
            const int ky = tap / 3, kx = tap - 3 * ky;
            float sy = (float)(py - 1 + ky) + dy;
            float sx = (float)(px - 1 + kx) + dx;
            const bool in = sy > 0.f && sy < (float)(p.Ho + 1) && sx > 0.f && sx < (float)(p.Wo + 1);
            const float msk = in ? 1.f / (1.f + __expf(-ml)) : 0.f;
            sy = in ? sy : (float)py;                          // weight 0: any row of the window will do
            sx = in ? sx : (float)px;
            const float yf = floorf(sy), xf = floorf(sx);
            const float ly_ = sy - yf, lx_ = sx - xf, hy = 1.f - ly_, hx = 1.f - lx_;
            const int Y0 = (int)yf, X0 = (int)xf;
            const int wy = Y0 - (y0 + 1 - h), wx = X0 - (x0 + 1 - h);
            const uint32_t gr0 = (uint32_t)((long long)p.x.guard + ((long long)b * Hp + Y0) * Wp + X0);
            uint4 o;
            if (wy >= 0 && wy <= LH - 2 && wx >= 0 && wx <= LW - 2) {
              const uint32_t l0 = (uint32_t)(wy * LW + wx), l2 = l0 + (uint32_t)LW;
              const uint32_t gr2 = gr0 + (uint32_t)Wp;
              o = make_uint4(xt_s + (l0 << 7) + ((gr0 & 7u) << 4), xt_s + ((l0 + 1u) << 7) + (((gr0 + 1u) & 7u) << 4),
                             xt_s + (l2 << 7) + ((gr2 & 7u) << 4), xt_s + ((l2 + 1u) << 7) + (((gr2 + 1u) & 7u) << 4));
            } else {
              o = make_uint4(0x80000000u | gr0, 0u, 0u, 0u);
            }
            tab_o[e] = o;
            tab_w[e] = make_float4(msk * hy * hx, msk * hy * lx_, msk * ly_ * hx, msk * ly_ * lx_);
          }
        }
        __syncwarp();
        mbar_wait(xbar, xph);
        xph ^= 1u;
        const unsigned char* xk = p.x.base + (((size_t)(p.x.chunk0 + kc) * p.x.rows) << 7);
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&empty[pst], pph);
          const uint32_t sA_s = smem_u32(ring + (size_t)pst * stage_bytes);
          if (!(p.dbg & 1)) {
            uint4 v[RPL][4][NS];
            float4 w[RPL];
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
              const uint4 o4 = lds128(tab_s + (uint32_t)(tap * TM + rr[i]) * 16u);
              const uint4 w4 = lds128(tab_s + (uint32_t)(9 * TM + tap * TM + rr[i]) * 16u);
              w[i] = make_float4(__uint_as_float(w4.x), __uint_as_float(w4.y), __uint_as_float(w4.z), __uint_as_float(w4.w));
              if (o4.x & 0x80000000u) {
                // outside the staged window: corner rows from global memory
                const uint32_t r0 = o4.x & 0x7fffffffu;
                const uint32_t rws[4] = {r0, r0 + 1u, r0 + (uint32_t)Wp, r0 + (uint32_t)Wp + 1u};
#pragma unroll
                for (int cn = 0; cn < 4; ++cn) {
                  const unsigned char* src = xk + ((size_t)rws[cn] << 7) + ((((rws[cn] & 7u) << 4)) ^ gx);
                  v[i][cn][0] = __ldg(reinterpret_cast<const uint4*>(src));
                  if (NS == 2) v[i][cn][NS - 1] = __ldg(reinterpret_cast<const uint4*>(src + plane_bytes));
                }
              } else {
                const uint32_t off[4] = {o4.x ^ gx, o4.y ^ gx, o4.z ^ gx, o4.w ^ gx};
#pragma unroll
                for (int cn = 0; cn < 4; ++cn) {
                  v[i][cn][0] = lds128(off[cn]);
                  if (NS == 2) v[i][cn][NS - 1] = lds128(off[cn] + xt_plane);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
              const float wc[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
              uint4 e0, e1 = make_uint4(0, 0, 0, 0);
              if (NS == 2) {
                // hi plane blended in fp32; lo plane (a 2^-11-scaled correction) in packed fp16 (2^-21 relative)
                float H[8];
                __half2 L[4];
#pragma unroll
                for (int cn = 0; cn < 4; ++cn) {
                  const uint32_t hv[4] = {v[i][cn][0].x, v[i][cn][0].y, v[i][cn][0].z, v[i][cn][0].w};
                  const uint32_t lv[4] = {v[i][cn][NS - 1].x, v[i][cn][NS - 1].y, v[i][cn][NS - 1].z, v[i][cn][NS - 1].w};
                  const __half2 wh = __float2half2_rn(wc[cn]);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 hf = unpack_h2(hv[j]);
                    const __half2 lh = *reinterpret_cast<const __half2*>(&lv[j]);
                    if (cn == 0) {
                      H[2 * j] = wc[0] * hf.x; H[2 * j + 1] = wc[0] * hf.y;
                      L[j] = __hmul2(wh, lh);
                    } else {
                      H[2 * j] = fmaf(wc[cn], hf.x, H[2 * j]); H[2 * j + 1] = fmaf(wc[cn], hf.y, H[2 * j + 1]);
                      L[j] = __hfma2(wh, lh, L[j]);
                    }
                  }
                }
                uint32_t eh[4], el[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 lf = __half22float2(L[j]);
                  const float a = fmaf(lf.x, LO_INV, H[2 * j]), b2 = fmaf(lf.y, LO_INV, H[2 * j + 1]);
                  const __half2 hh = __floats2half2_rn(a, b2);
                  const float2 hb = __half22float2(hh);
                  eh[j] = *reinterpret_cast<const uint32_t*>(&hh);
                  el[j] = pack_h2((a - hb.x) * LO_SCALE, (b2 - hb.y) * LO_SCALE);
                }
                e0 = make_uint4(eh[0], eh[1], eh[2], eh[3]);
                e1 = make_uint4(el[0], el[1], el[2], el[3]);
              } else {
                // bf16 mode: the blend itself runs in packed bf16 (four HFMA2.BF16 per corner, no unpack / pack)
                __nv_bfloat162 o2[4];
#pragma unroll
                for (int cn = 0; cn < 4; ++cn) {
                  const uint32_t bv[4] = {v[i][cn][0].x, v[i][cn][0].y, v[i][cn][0].z, v[i][cn][0].w};
                  const __nv_bfloat162 wb = __float2bfloat162_rn(wc[cn]);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(&bv[j]);
                    o2[j] = cn == 0 ? __hmul2(wb, xv) : __hfma2(wb, xv, o2[j]);
                  }
                }
                e0 = make_uint4(*reinterpret_cast<uint32_t*>(&o2[0]), *reinterpret_cast<uint32_t*>(&o2[1]),
                                *reinterpret_cast<uint32_t*>(&o2[2]), *reinterpret_cast<uint32_t*>(&o2[3]));
              }
              sts128(sA_s + soff[i], e0);
              if (NS == 2) sts128(sA_s + a_plane + soff[i], e1);
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[pst]);
          if (++pst == p.SA) { pst = 0; pph ^= 1u; }
        }
      }
    }
  } else if (warp < D_PROD_WARPS + 4) {
    // ==================================================================== epilogue warps
    if (WIDE) reg_inc<152>();
    else reg_inc<120>();
    const int q = warp & 3;
    if constexpr (NS == 2) {
      constexpr int BAND_KB = BAND_KSTEPS / 4;
      fp32_epilogue<1, true, WIDE ? 8 : 4>(p, tmem, total, (nkb + BAND_KB - 1) / BAND_KB, q, 0, lane, slot_full, slot_empty, acc_empty, epi_ss);
    } else {
      uint32_t as = 0, accph = 0, ss_buf = 0;
      int ss_nt = -1;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const uint32_t ss = epi_stage_ss<1>(p, 0, ss_nt, ss_buf, epi_ss, q * 32 + lane);      // n_tiles == 1
        mbar_wait_backoff(&acc_full[as], accph);
        tc_fence_after();
        if (!(p.dbg & 4))
          epilogue_tile<NS, 1>(p, tmem + as * (uint32_t)NT + ((uint32_t)(q * 32) << 16), 0, 0, q * 32 + lane, 0u, 0, nullptr, ss,
                               tile_row_m(p, t, q * 32 + lane));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        as ^= 1u;
        if (as == 0) accph ^= 1u;
      }
    }
  } else {
    reg_dec<40>();
    if (warp == 20) {
      // ==================================================================== B loader
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 1;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
          int kc = 0, tap = 0;
          for (int kb = 0; kb < nkb; ++kb) {
            const int blk = tap * p.KC + kc;                   // producers walk (kc outer, tap inner); packed blocks are (tap, kc)
            mbar_wait(&empty[s], ph);
            mbar_arrive_expect_tx(&full[s], b_bytes);
            bulk_g2s(ring + (size_t)s * stage_bytes + a_bytes, p.wpack + (size_t)blk * b_bytes, b_bytes, &full[s]);
            if (++s == p.SA) { s = 0; ph ^= 1u; }
            if (++tap == 9) { tap = 0; ++kc; }
          }
        }
      }
    } else if (warp == 21) {
      // ==================================================================== MMA issuer
      const uint32_t idesc = NS == 2 ? idesc_f16_f32(TM, NT) : idesc_bf16_f32(TM, NT);
      const uint32_t idesc_w = idesc_f16_f32(TM, 2 * NT);
      constexpr int BAND_KB = BAND_KSTEPS / 4;
      int s = 0;
      uint32_t ph = 0, as = 0, accph = 1;
      Ring rs{0u, 1u, p.nslots}, rd{0u, 1u, p.nd1};
      const uint32_t smem0 = smem_u32(ring);
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        uint32_t tacc = 0, tD0 = 0, tD1 = 0, sl = 0;
        if constexpr (NS == 2) {
          if (!p.wide) {
            mbar_wait(&acc_empty[rd.i], rd.ph);
            tD1 = tmem + (uint32_t)(p.nslots + (int)rd.i) * (uint32_t)NT;
            rd.next();
          }
        } else {
          mbar_wait(&acc_empty[as], accph);
          tacc = tmem + as * (uint32_t)NT;
        }
        int kin = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[s], ph);
          const uint32_t a0 = smem0 + (uint32_t)s * stage_bytes, b0 = a0 + a_bytes;
          if constexpr (NS == 2) {
            if (kin == 0) {
              sl = rs.i;
              mbar_wait(&slot_empty[sl], rs.ph);
              tD0 = tmem + sl * (uint32_t)p.slot_cols;
              rs.next();
            }
            tc_fence_after();
            if (elect_one()) {
              if (p.wide) issue_kblock_f32w(a0, a_plane, b0, tD0, (uint32_t)NT, idesc_w, idesc, kin ? 1u : 0u);
              else issue_kblock_f32(a0, a_plane, b0, b_plane, tD0, tD1, idesc, kin ? 1u : 0u, kb ? 1u : 0u);
              mma_commit(&empty[s]);
              if (kin + 1 == BAND_KB || kb == nkb - 1) mma_commit(&slot_full[sl]);
            }
            __syncwarp();
            if (++kin == BAND_KB || kb == nkb - 1) kin = 0;
          } else {
            tc_fence_after();
            if (elect_one()) {
              if (kb == 0) issue_kblock<NS, true>(a0, a_plane, b0, b_plane, tacc, (uint32_t)NT, idesc, 0);
              else issue_kblock<NS, false>(a0, a_plane, b0, b_plane, tacc, (uint32_t)NT, idesc, 0);
              mma_commit(&empty[s]);
            }
            __syncwarp();
          }
          if (++s == p.SA) { s = 0; ph ^= 1u; }
        }
        if constexpr (NS != 2) {
          if (elect_one()) mma_commit(&acc_full[as]);
          __syncwarp();
          as ^= 1u;
          if (as == 0) accph ^= 1u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 21) tmem_dealloc_dyn(tmem, p.tmem_cols);
}

// =============================================================================== weights
// Wm [Cout][Kpad] fp32 -> wpack[n tile][K block][plane][NT x 64 SWIZZLE_128B image]
__global__ void pack_weight_planes_kernel(const float* __restrict__ wm, unsigned char* __restrict__ wpack, int Cout,
                                          int Kpad, int NT, int NS) {
  const int nkb = Kpad / 64;
  const long long total = (long long)Cout * Kpad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % Kpad), n = (int)(e / Kpad);
    const int kb = k / 64, kk = k % 64, t = n / NT, row = n % NT;
    const float v = wm[e];
    const size_t base = ((size_t)t * nkb + kb) * (size_t)NT * 128 * NS;
    const size_t off = sw128_offset(row, kk / 8) + (kk % 8) * 2;
    if (NS == 2) {
      const float c = fminf(fmaxf(v, -65504.f), 65504.f);
      const __half hi = __float2half_rn(c);
      *reinterpret_cast<__half*>(wpack + base + off) = hi;
      *reinterpret_cast<__half*>(wpack + base + (size_t)NT * 128 + off) = __float2half_rn((c - __half2float(hi)) * LO_SCALE);
    } else {
      *reinterpret_cast<__nv_bfloat16*>(wpack + base + off) = __float2bfloat16_rn(v);
    }
  }
}

static int pick_ntile(int Cout, int NS) {
  if (Cout < 16 || Cout % 16) return -1;
  const int cap = (g_dbg & 64) ? 64 : NS == 2 ? 128 : 256;
  int nt = Cout;
  if (nt > cap) {
    nt = cap;
    while (Cout % nt) nt -= 16;
  }
  return nt;
}
// accumulator plan: R hi*hi accumulators (+ D1) per stage; two stages when they fit in 512 columns
// short_k: the truncation bias grows with the number of K steps (tools/acc_probe.py); a K <= 576 sum
// stays below 4e-6 relative without rotation, and R = 1 lets a 128-wide N tile keep TWO accumulator
// stages so the epilogue overlaps the next tile's MMAs (the 64 -> 768 head convolution)
static void plan_acc(ConvP& p, int NS, bool short_k = false) {
  (void)short_k;
  p.acc_r = 1;
  p.acc_stages = 2;
  int need = 2 * p.NT;
  if (NS == 2) {
    // fp32 mode (band-drain accumulation): D0 band slots + D1 buffers, NT <= 128 columns each.  The MMA warp can
    // run as many bands ahead of the epilogue warps as there are slots, which is what hides the previous tile's
    // scale / activation / store phase: as many as the 512 columns give (3 + 1 for 128-wide tiles, 6 + 2 below)
    // Wide form (issue_kblock_f32w): slots of [D0 | D1], no D1 buffers -- 4 slots at NT = 64, 6 below, 2 at 128.
    // Measured (tools/wide_bench.py): 64 -> 64 3x3 163 -> 147 us, offset / mask convolutions (NT = 32) 79 -> 66 us;
    // 128-wide tiles are left with 2 slots: 1-2 % faster on long K (128 -> 128 3x3 130.7 -> 129.0 us), slower on the
    // K = 576 head convolution (710 -> 730 us), and their D1 would be rounded into the sum once per band instead of
    // once per tile -- they keep the three-MMA form.  Debug flag 16384 disables the wide form, 32768 forces it everywhere.
    p.wide = !(p.dbg & 16384) && (p.NT <= 64 || (p.dbg & 32768));
    if (p.wide) {
      p.slot_cols = 2 * p.NT;
      p.nd1 = 0;
      p.nslots = 512 / p.slot_cols > 6 ? 6 : 512 / p.slot_cols;
      need = p.nslots * p.slot_cols;
    } else {
      const int units = 512 / p.NT;
      p.slot_cols = p.NT;
      p.nd1 = units >= 8 ? 2 : 1;
      p.nslots = units - p.nd1 > 6 ? 6 : units - p.nd1;
      need = (p.nslots + p.nd1) * p.NT;
    }
  }
  int c = 32;
  while (c < need) c <<= 1;
  p.tmem_cols = c;
}

template <int NS>
static int launch_shift(ConvP& p, cudaStream_t st) {
  const int a_stage = p.a_rows * 128 * NS, b_stage = p.NT * 128 * NS;
  // a 3x3 band consumes THREE B slots before it releases any (SB < 3 deadlocks); one more slot is the
  // whole weight prefetch, so the epilogue's store staging is only taken when SB >= 4 still fits
  const int min_sb = p.taps == 9 ? 3 : 2;
  int fixed = 0, SA = 0, SB = 0;
  // Resident weights: when one N tile covers Cout and all K blocks fit next to >= 2 A stages, the B ring gets one slot
  // per K block and is filled once per CTA instead of once per tile (64 -> 64 3x3: 147 KB of weights per 128-row tile
  // otherwise -- more L2 -> SM traffic than the A bands).  Debug flag 2048 disables it.
  p.b_resident = 0;
  if (p.n_tiles == 1 && p.nkb <= S_MAX_SB && p.nkb >= min_sb && !(p.dbg & 2048)) {
    const int stg = (NS == 1 && p.epi == SGTA_EPI_PL && !(p.dbg & 8)) ? 16384 * NS : 0;
    const int fx = 1024 + 1024 + EPI_SS_BYTES + p.heads_bytes + stg;
    if (2 * a_stage + p.nkb * b_stage + fx <= SMEM_LIMIT) {
      p.b_resident = 1;
      p.stg_bytes = stg;
      fixed = fx;
      SB = p.nkb;
      SA = (SMEM_LIMIT - fx - SB * b_stage) / a_stage;
      if (SA > 6) SA = 6;
    }
  }
  // (fp32 mode: measured no gain -- the hi/lo tile needs 32 KB that the weight ring uses better)
  for (int with_stg = p.b_resident ? -1 : (NS == 1 && p.epi == SGTA_EPI_PL && !(p.dbg & 8)) ? 1 : 0; with_stg >= 0; --with_stg) {
    p.stg_bytes = with_stg ? 16384 * NS : 0;
    fixed = 1024 + 1024 + EPI_SS_BYTES + p.heads_bytes + p.stg_bytes;
    SB = 4; SA = 3;
    while (SB > min_sb && SA * a_stage + SB * b_stage + fixed > SMEM_LIMIT) --SB;
    while (SA > 1 && SA * a_stage + SB * b_stage + fixed > SMEM_LIMIT) --SA;
    const bool fits = SA * a_stage + SB * b_stage + fixed <= SMEM_LIMIT;
    if (fits && (!with_stg || (SB >= min_sb + 1 && SA >= 2))) break;
    if (!with_stg) { set_error("conv_shift: tile does not fit shared memory"); return SGTA_EUNSUPPORTED; }
  }
  // spend what is left on deeper rings
  if (!p.b_resident) {
    while (SA < 6 && (SA + 1) * a_stage + SB * b_stage + fixed <= SMEM_LIMIT && SA <= SB) ++SA;
    while (SB < 8 && SA * a_stage + (SB + 1) * b_stage + fixed <= SMEM_LIMIT) ++SB;
  }
  p.SA = SA; p.SB = SB;
  const int smem = SA * a_stage + SB * b_stage + fixed;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int total = p.m_major ? p.m_tiles : p.m_tiles * p.n_tiles;
  const int grid = total < sms ? total : sms;
  if (p.spk) {
    cudaFuncSetAttribute(conv_shift_kernel<NS, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_k(conv_shift_kernel<NS, 1, 1>, grid, S_THREADS, smem, st, p);
  } else {
    cudaFuncSetAttribute(conv_shift_kernel<NS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_k(conv_shift_kernel<NS, 1>, grid, S_THREADS, smem, st, p);
  }
  return check_launch("conv_shift_kernel");
}

template <int PROD, int NS>
static int launch_gather(ConvP& p, cudaStream_t st) {
  const int a_bytes = TM * 128 * NS, b_bytes = p.NT * 128 * NS;
  p.stg_bytes = NS == 1 && p.epi == SGTA_EPI_PL && !(p.dbg & 8) ? 16384 * NS : 0;
  const int fixed = 1024 + 512 + EPI_SS_BYTES + p.stg_bytes + (PROD == PROD_DCN ? 9 * TM * 32 : 0);
  // weights resident in shared memory when one CTA sees a single N tile and they leave room for
  // >= 2 A stages: no per-K-block weight copy (a single thread's bulk copies serialise, ~530 clk each)
  p.b_resident = p.n_tiles == 1 && fixed + p.nkb * b_bytes + 2 * a_bytes <= SMEM_LIMIT;
  const int stage = p.b_resident ? a_bytes : a_bytes + b_bytes;
  const int avail = SMEM_LIMIT - fixed - (p.b_resident ? p.nkb * b_bytes : 0);
  int SA = avail / stage;
  // the gathers live on L1 hits (neighbouring taps / corners touch the same lines): keep the
  // operand ring short so the unified L1 / shared carve-out leaves a large cache
  // (consuming G > 1 stages per issue group helps the MMA warp of the small-N layers but the
  // deeper ring it needs evicts the L1 the producers live on: measured slower, so G stays 1)
  if (SA > 3) SA = 3;                       // (DCN with 16 producer warps: 2 -> 3.17 ms, 3 -> 3.02 ms, 4 -> 3.06 ms per step)
  // (2-D tiles with SA = 2, i.e. a 92 KB L1: within 2 % of SA = 3 on every DCN shape)
  if (PROD == PROD_DCN && (p.dbg >> 20) & 7) { const int want = (p.dbg >> 20) & 7; if (want >= 2 && want <= avail / stage) SA = want; }   // experiments
  if (SA < 2) { set_error("conv_gather: tile does not fit shared memory"); return SGTA_EUNSUPPORTED; }
  p.SA = SA; p.SB = 0;
  const int smem = SA * stage + fixed + (p.b_resident ? p.nkb * b_bytes : 0);
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int total = p.tile2d ? p.tiles_x * p.tiles_y * p.x.B : p.m_tiles * p.n_tiles;
  const int grid = total < sms ? total : sms;
  if (NS == 2 && PROD == PROD_DCN && p.NT > 64) {
    constexpr int W = (NS == 2 && PROD == PROD_DCN) ? 1 : 0;     // (only this combination instantiates the WIDE variant)
    cudaFuncSetAttribute(conv_gather_kernel<PROD, NS, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_k(conv_gather_kernel<PROD, NS, W>, grid, GThreads<PROD>::value, smem, st, p);
  } else {
    cudaFuncSetAttribute(conv_gather_kernel<PROD, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_k(conv_gather_kernel<PROD, NS>, grid, GThreads<PROD>::value, smem, st, p);
  }
  return check_launch("conv_gather_kernel");
}

// dcn_tile_kernel when the map tiles into 8 x 16 blocks, one N tile covers Cout and the staged window fits
template <int NS>
static int try_launch_dcn_tile(ConvP& p, cudaStream_t st) {
  if ((p.dbg & 32) || p.Ho % 8 || p.Wo % 16 || p.n_tiles != 1) return -1;
  const int a_bytes = TM * 128 * NS, b_bytes = p.NT * 128 * NS;
  const int fixed = 1024 + 512 + EPI_SS_BYTES + 9 * TM * 32;
  for (int SA = NS == 2 ? 2 : 3; SA >= 2; --SA) {
    for (int h = 3; h >= 2; --h) {
      const int LH = 8 + 2 * h, LW = 16 + 2 * h;
      const int xt_plane = (LH * LW * 128 + 1023) & ~1023;
      const int smem = xt_plane * NS + SA * (a_bytes + b_bytes) + fixed;
      if (smem > SMEM_LIMIT) continue;
      p.tile2d = 1; p.tiles_x = p.Wo / 16; p.tiles_y = p.Ho / 8; p.halo = h; p.LW = LW; p.LH = LH; p.xt_plane = xt_plane;
      p.SA = SA; p.SB = 0; p.stg_bytes = 0; p.b_resident = 0;
      static int sms = 0;
      if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
      const int total = p.tiles_x * p.tiles_y * p.x.B;
      const int grid = total < sms ? total : sms;
      if (NS == 2 && p.NT > 64) {
        cudaFuncSetAttribute(dcn_tile_kernel<NS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        launch_k(dcn_tile_kernel<NS, 1>, grid, D_THREADS, smem, st, p);
      } else {
        cudaFuncSetAttribute(dcn_tile_kernel<NS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        launch_k(dcn_tile_kernel<NS, 0>, grid, D_THREADS, smem, st, p);
      }
      return check_launch("dcn_tile_kernel");
    }
  }
  return -1;
}

static bool view_ok(const sgta_planes* v, int layout) {
  return v && v->data && v->layout == layout && (v->nplanes == 1 || v->nplanes == 2) && v->B > 0 && v->H > 0 &&
         v->W > 0 && v->guard >= 0 && v->rows >= (int64_t)v->guard + (int64_t)v->B * (v->H + 2 * v->border) * (v->W + 2 * v->border);
}

}  // namespace sgta

using namespace sgta;

extern "C" int sgta_planes_ntile(int Cout, int nplanes) { return pick_ntile(Cout, nplanes); }

extern "C" int sgta_debug_flags(int flags) { int old = g_dbg; g_dbg = flags; return old; }

extern "C" int sgta_planes_guard(int W) { return ((W + 2) * 8 + 160 + 7) & ~7; }

extern "C" int64_t sgta_planes_wpack_bytes(int Cout, int Kpad, int nplanes) {
  if (pick_ntile(Cout, nplanes) < 0 || Kpad <= 0 || Kpad % 64) return -1;
  return (int64_t)Cout * Kpad * 2 * nplanes;
}

extern "C" int sgta_planes_pack_weight(const void* wm_f32, void* wpack, int Cout, int Kpad, int nplanes, void* stream) {
  SGTA_REQUIRE(wm_f32 && wpack, "sgta_planes_pack_weight: null pointer");
  SGTA_REQUIRE(nplanes == 1 || nplanes == 2, "sgta_planes_pack_weight: nplanes must be 1 or 2");
  const int nt = pick_ntile(Cout, nplanes);
  SGTA_REQUIRE(nt > 0 && Kpad > 0 && Kpad % 64 == 0, "sgta_planes_pack_weight: need Cout %% 16 == 0 and Kpad %% 64 == 0 (got %d, %d)", Cout, Kpad);
  const long long total = (long long)Cout * Kpad;
  const int blocks = cdiv(total, 256) > 2368 ? 2368 : cdiv(total, 256);
  pack_weight_planes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)wm_f32, (unsigned char*)wpack, Cout, Kpad, nt, nplanes);
  return check_launch("pack_weight_planes_kernel");
}

static int fill_output(ConvP& p, const sgta_planes* y, void* y_f32, int64_t ld_f32, int Cout, int epi, int n_valid,
                       int B, int Ho, int Wo, int nplanes, const char* who) {
  p.epi = epi;
  p.n_valid = n_valid > 0 ? n_valid : Cout;
  if (epi == SGTA_EPI_PL) {
    SGTA_REQUIRE(view_ok(y, SGTA_LAYOUT_PL) && y->nplanes == nplanes && y->border == 1, "%s: bad PL output view", who);
    SGTA_REQUIRE(Cout % 64 == 0 && y->chunk0 + Cout / 64 <= y->nchunks, "%s: PL output needs Cout %% 64 == 0 inside the buffer", who);
  } else if (epi == SGTA_EPI_SC || epi == SGTA_EPI_STEM) {
    SGTA_REQUIRE(view_ok(y, SGTA_LAYOUT_SC) && y->nplanes == nplanes && y->border == 1, "%s: bad SC output view", who);
    SGTA_REQUIRE(epi == SGTA_EPI_STEM ? (Cout == 32 && y->nchunks == 16) : y->nchunks == Cout, "%s: SC output channel mismatch", who);
  } else if (epi == SGTA_EPI_STEM_SP) {
    SGTA_REQUIRE(view_ok(y, SGTA_LAYOUT_PL) && y->nplanes == nplanes && y->border == 1 && Cout == 128 && y->chunk0 < y->nchunks,
                 "%s: the super-pixel stem writes a 64-channel PL view (Cout = 4 pixels x 32)", who);
  } else if (epi == SGTA_EPI_SP2SC) {
    SGTA_REQUIRE(view_ok(y, SGTA_LAYOUT_SC) && y->nplanes == nplanes && y->border == 1 && Cout == 64 && y->nchunks == 16,
                 "%s: EPI_SP2SC writes 4 pixels x 16 channels per row into a 16-channel SC view", who);
    SGTA_REQUIRE(y->B == B && y->H == Ho && y->W == 4 * Wo, "%s: output geometry mismatch", who);
    p.y = make_view(y);
  } else {
    SGTA_REQUIRE(epi == SGTA_EPI_F32ROWS || epi == SGTA_EPI_NCHW, "%s: bad epilogue", who);
    SGTA_REQUIRE(y_f32 && (epi == SGTA_EPI_NCHW || ld_f32 >= Cout), "%s: bad fp32 output", who);
  }
  if (epi == SGTA_EPI_PL || epi == SGTA_EPI_SC || epi == SGTA_EPI_STEM || epi == SGTA_EPI_STEM_SP) {
    SGTA_REQUIRE(y->B == B && y->H == Ho && y->W == Wo, "%s: output geometry mismatch", who);
    p.y = make_view(y);
  }
  p.yf = (float*)y_f32;
  p.ldyf = ld_f32;
  return SGTA_OK;
}

extern "C" int sgta_planes_conv(const sgta_planes* x, const void* wpack, const void* scale, const void* shift,
                                const sgta_planes* res, const sgta_planes* y, void* y_f32, int64_t ld_f32, int Cin,
                                int Cout, int ksize, int stride, int act, int epi, int n_valid, void* stream) {
  SGTA_REQUIRE(x && wpack && scale && shift, "sgta_planes_conv: null pointer");
  SGTA_REQUIRE(x->nplanes == 1 || x->nplanes == 2, "sgta_planes_conv: nplanes must be 1 or 2");
  const int NS = x->nplanes;
  const int NT = pick_ntile(Cout, NS);
  SGTA_REQUIRE(NT > 0, "sgta_planes_conv: need Cout %% 16 == 0 (got %d)", Cout);
  SGTA_REQUIRE(stride == 1 || stride == 2, "sgta_planes_conv: stride must be 1 or 2");
  ConvP p{};
  p.dbg = g_dbg;
  p.x = make_view(x);
  p.wpack = (const unsigned char*)wpack; p.scale = (const float*)scale; p.shift = (const float*)shift;
  p.act = act; p.stride = stride; p.sx = stride; p.NT = NT; p.n_tiles = Cout / NT;
  const int pad = ksize / 2;
  const int Ho = (x->H + 2 * pad - ksize) / stride + 1, Wo = (x->W + 2 * pad - ksize) / stride + 1;
  SGTA_REQUIRE(Ho > 0 && Wo > 0, "sgta_planes_conv: empty output");
  const long long P = (long long)x->B * (Ho + 2) * (Wo + 2);
  SGTA_REQUIRE(P < (1ll << 31) - 2 * TM, "sgta_planes_conv: too many pixels");
  p.P = (int)P; p.Ho = Ho; p.Wo = Wo; p.m_tiles = cdiv(P, TM);
  p.magic_wp = magic_u64(Wo + 2); p.magic_hp = magic_u64(Ho + 2);
  int rc = fill_output(p, y, y_f32, ld_f32, Cout, epi, n_valid, x->B, Ho, Wo, NS, "sgta_planes_conv");
  if (rc) return rc;
  if (res) {
    SGTA_REQUIRE(view_ok(res, SGTA_LAYOUT_PL) && res->nplanes == NS && res->B == x->B && res->H == Ho && res->W == Wo &&
                 res->border == 1 && epi == SGTA_EPI_PL, "sgta_planes_conv: bad residual view");
    p.res = make_view(res);
  }
  p.in_Wp = x->W + 2 * x->border; p.in_Hp = x->H + 2 * x->border;
  cudaStream_t st = (cudaStream_t)stream;
  if (x->layout == SGTA_LAYOUT_PL) {
    SGTA_REQUIRE(view_ok(x, SGTA_LAYOUT_PL) && x->border == 1, "sgta_planes_conv: bad PL input view");
    SGTA_REQUIRE(Cin % 64 == 0 && x->chunk0 + Cin / 64 <= x->nchunks, "sgta_planes_conv: PL input needs Cin %% 64 == 0 inside the buffer");
    p.KC = Cin / 64;
    if (stride == 1) {
      SGTA_REQUIRE(ksize == 1 || ksize == 3, "sgta_planes_conv: PL stride-1 kernels are 1x1 or 3x3");
      p.taps = ksize * ksize; p.nkb = p.taps * p.KC; p.a_rows = ksize == 3 ? 144 : 136;
      // EPI_SP2SC = level0 over super-pixels: the weights come from planes.superpixel_weight (gin = gout = 4, 16
      // channels), whose neighbour taps are zero outside one pixel -- those K steps are not issued
      p.spk = epi == SGTA_EPI_SP2SC && ksize == 3 && Cin == 64 && !(p.dbg & 512);
      plan_acc(p, NS);
      const int Wp = x->W + 2;
      SGTA_REQUIRE(x->guard >= Wp + 1 + 8 && x->rows >= (int64_t)x->guard + (int64_t)p.m_tiles * TM + Wp + 16 + 8,
                   "sgta_planes_conv: input guard rows too small for the halo");
      return NS == 2 ? launch_shift<2>(p, st) : launch_shift<1>(p, st);
    }
    SGTA_REQUIRE(ksize == 3, "sgta_planes_conv: PL strided kernels are 3x3");
    p.taps = 9; p.nkb = 9 * p.KC;
    plan_acc(p, NS);
    return NS == 2 ? launch_gather<PROD_STRIDE, 2>(p, st) : launch_gather<PROD_STRIDE, 1>(p, st);
  }
  // SC input: K blocks described by the caller through sgta_planes_conv_sc
  set_error("sgta_planes_conv: SC inputs go through sgta_planes_conv_sc");
  return SGTA_EINVAL;
}

extern "C" int sgta_planes_conv_heads(const sgta_planes* x, const void* wpack, const void* scale, const void* shift,
                                      const void* w2, const void* b2, void* const* out, const int* nout, int n_heads,
                                      int sigmoid_mask, int Cin, int hid, void* stream) {
  SGTA_REQUIRE(x && wpack && scale && shift && w2 && b2 && out && nout, "sgta_planes_conv_heads: null pointer");
  SGTA_REQUIRE(view_ok(x, SGTA_LAYOUT_PL) && x->border == 1 && x->nplanes == 2, "sgta_planes_conv_heads: fp32-mode PL input view expected");
  SGTA_REQUIRE(n_heads >= 1 && n_heads <= 4 && Cin % 64 == 0 && x->chunk0 + Cin / 64 <= x->nchunks, "sgta_planes_conv_heads: bad shape");
  const int NS = 2, Cout = n_heads * hid;
  const int NT = pick_ntile(Cout, NS);
  SGTA_REQUIRE(NT > 0 && hid % NT == 0, "sgta_planes_conv_heads: hid (%d) must be a multiple of the N tile (%d)", hid, NT);
  ConvP p{};
  p.dbg = g_dbg;
  p.x = make_view(x);
  p.wpack = (const unsigned char*)wpack; p.scale = (const float*)scale; p.shift = (const float*)shift;
  p.act = SGTA_ACT_RELU; p.stride = 1; p.sx = 1; p.NT = NT; p.n_tiles = Cout / NT;
  const long long P = (long long)x->B * (x->H + 2) * (x->W + 2);
  SGTA_REQUIRE(P < (1ll << 31) - 2 * TM, "sgta_planes_conv_heads: too many pixels");
  p.P = (int)P; p.Ho = x->H; p.Wo = x->W; p.m_tiles = cdiv(P, TM);
  p.magic_wp = magic_u64(x->W + 2); p.magic_hp = magic_u64(x->H + 2);
  p.epi = SGTA_EPI_HEADS; p.n_valid = Cout;
  p.in_Wp = x->W + 2; p.in_Hp = x->H + 2;
  p.KC = Cin / 64; p.taps = 9; p.nkb = 9 * p.KC; p.a_rows = 144;
  p.m_major = 1; p.n_heads = n_heads; p.h_tiles = hid / NT; p.sig_mask = sigmoid_mask;
  p.w2 = (const float*)w2; p.b2 = (const float*)b2;
  int off = 0;
  for (int h = 0; h < n_heads; ++h) {
    SGTA_REQUIRE(out[h] && nout[h] >= 1 && nout[h] <= 8, "sgta_planes_conv_heads: head %d: bad output", h);
    p.out2[h] = (float*)out[h]; p.nout[h] = nout[h];
    p.w2_stride[h] = nout[h] <= 2 ? 2 : nout[h] <= 4 ? 4 : 8;
    p.w2_off[h] = off;
    off += hid * p.w2_stride[h];
  }
  p.w2_floats = off;
  p.heads_bytes = off * 4 + TM * 8 * 4;                      // fused weights + the half-to-half exchange buffer
  plan_acc(p, NS);
  const int Wp = x->W + 2;
  SGTA_REQUIRE(x->guard >= Wp + 1 + 8 && x->rows >= (int64_t)x->guard + (int64_t)p.m_tiles * TM + Wp + 16 + 8,
               "sgta_planes_conv_heads: input guard rows too small for the halo");
  return launch_shift<2>(p, (cudaStream_t)stream);
}

extern "C" int sgta_planes_conv_sc(const sgta_planes* x, const void* wpack, const void* scale, const void* shift,
                                   const sgta_planes* y, int Cout, int stride, int stride_x, int Ho, int Wo, int nkb,
                                   int seg_groups, const int* seg_off /*HOST [nkb*2]*/, int act, int epi, void* stream) {
  SGTA_REQUIRE(x && wpack && scale && shift && y && seg_off, "sgta_planes_conv_sc: null pointer");
  SGTA_REQUIRE(view_ok(x, SGTA_LAYOUT_SC), "sgta_planes_conv_sc: bad SC input view");
  const int NS = x->nplanes, C = x->nchunks;
  SGTA_REQUIRE(C == 4 || C == 16 || C == 32, "sgta_planes_conv_sc: C must be 4, 16 or 32 (got %d)", C);
  SGTA_REQUIRE(nkb >= 1 && nkb <= 8 && (seg_groups == 8 || seg_groups == 4), "sgta_planes_conv_sc: bad K-block table");
  const int NT = pick_ntile(Cout, NS);
  SGTA_REQUIRE(NT > 0, "sgta_planes_conv_sc: need Cout %% 16 == 0 (got %d)", Cout);
  ConvP p{};
  p.dbg = g_dbg;
  p.x = make_view(x);
  p.wpack = (const unsigned char*)wpack; p.scale = (const float*)scale; p.shift = (const float*)shift;
  SGTA_REQUIRE(stride >= 1 && stride_x >= 1, "sgta_planes_conv_sc: strides must be positive");
  p.act = act; p.stride = stride; p.sx = stride_x; p.NT = NT; p.n_tiles = Cout / NT; p.nkb = nkb; plan_acc(p, NS);
  const long long P = (long long)x->B * (Ho + 2) * (Wo + 2);
  SGTA_REQUIRE(P < (1ll << 31) - 2 * TM, "sgta_planes_conv_sc: too many pixels");
  p.P = (int)P; p.Ho = Ho; p.Wo = Wo; p.m_tiles = cdiv(P, TM);
  p.magic_wp = magic_u64(Wo + 2); p.magic_hp = magic_u64(Ho + 2);
  int rc = fill_output(p, y, nullptr, 0, Cout, epi, 0, x->B, Ho, Wo, NS, "sgta_planes_conv_sc");
  if (rc) return rc;
  p.in_Wp = x->W + 2 * x->border; p.in_Hp = x->H + 2 * x->border;
  p.nkb = nkb; p.seg_groups = seg_groups; p.vec8 = C == 4;
  for (int i = 0; i < nkb * 2; ++i) p.seg_off[i] = seg_off[i];
  cudaStream_t st = (cudaStream_t)stream;
  return NS == 2 ? launch_gather<PROD_SMALLC, 2>(p, st) : launch_gather<PROD_SMALLC, 1>(p, st);
}

extern "C" int sgta_planes_dcn(const sgta_planes* x, const void* offset_mask, const void* wpack, const void* scale,
                               const void* shift, const sgta_planes* y, int Cin, int Cout, int relu, void* stream) {
  SGTA_REQUIRE(x && offset_mask && wpack && scale && shift && y, "sgta_planes_dcn: null pointer");
  SGTA_REQUIRE(view_ok(x, SGTA_LAYOUT_PL) && x->border == 1, "sgta_planes_dcn: bad PL input view");
  const int NS = x->nplanes;
  SGTA_REQUIRE(Cin % 64 == 0 && x->chunk0 + Cin / 64 <= x->nchunks, "sgta_planes_dcn: need Cin %% 64 == 0 inside the buffer");
  const int NT = pick_ntile(Cout, NS);
  SGTA_REQUIRE(NT > 0 && Cout % 64 == 0, "sgta_planes_dcn: need Cout %% 64 == 0 (got %d)", Cout);
  ConvP p{};
  p.dbg = g_dbg;
  p.x = make_view(x);
  p.om = (const float*)offset_mask;
  p.wpack = (const unsigned char*)wpack; p.scale = (const float*)scale; p.shift = (const float*)shift;
  p.act = relu ? SGTA_ACT_RELU : SGTA_ACT_NONE; p.stride = 1; p.NT = NT; p.n_tiles = Cout / NT;
  p.KC = Cin / 64; p.nkb = 9 * p.KC;
  plan_acc(p, NS);
  const long long P = (long long)x->B * (x->H + 2) * (x->W + 2);
  SGTA_REQUIRE(P < (1ll << 31) - 2 * TM, "sgta_planes_dcn: too many pixels");
  p.P = (int)P; p.Ho = x->H; p.Wo = x->W; p.m_tiles = cdiv(P, TM);
  p.magic_wp = magic_u64(x->W + 2); p.magic_hp = magic_u64(x->H + 2);
  int rc = fill_output(p, y, nullptr, 0, Cout, SGTA_EPI_PL, 0, x->B, x->H, x->W, NS, "sgta_planes_dcn");
  if (rc) return rc;
  p.in_Wp = x->W + 2; p.in_Hp = x->H + 2;
  p.KC = Cin / 64; p.taps = 9; p.nkb = 9 * p.KC;
  cudaStream_t st = (cudaStream_t)stream;
  // window-staged kernel where it measures faster: bf16 mode (1.72 vs 1.87 ms per step); in fp32 mode the blend is
  // instruction-bound either way and the __ldg gather wins (3.55 vs 3.80 ms) -- debug flag 256 forces it, 32 disables it
  rc = -1;
  if (NS == 1 || (p.dbg & 256)) rc = NS == 2 ? try_launch_dcn_tile<2>(p, st) : try_launch_dcn_tile<1>(p, st);
  if (rc >= 0) return rc;
  p.tile2d = 0;
  // fp32 mode: the __ldg gather over 2-D tiles.  A tile of 128 consecutive rows is a 1.3-row strip of a 96-wide map whose
  // 9 x 4 corner reads span ~5 rows x 98 pixels x 256 B = 125 KB -- twice the L1 that the 190 KB operand ring leaves
  // (ncu: L1 hit rate 48 %); an 8 x 16 block touches ~12 x 20 pixels = 61 KB.  Same per-pixel arithmetic in the same
  // order: the outputs are bit-identical to the strip tiling (debug flag 524288 keeps that one).  Measured: 128 -> 64
  // @48x48 157 -> 129 us, 128 -> 128 189 -> 155 us, 64 -> 64 @96x96 303 -> 294 us; DCN per step 3.10 -> 2.94 ms.
  if (NS == 2 && p.n_tiles == 1 && p.Ho % 8 == 0 && p.Wo % 16 == 0 && !(p.dbg & 524288)) {
    p.tile2d = 1; p.tiles_x = p.Wo / 16; p.tiles_y = p.Ho / 8;
  }
  return NS == 2 ? launch_gather<PROD_DCN, 2>(p, st) : launch_gather<PROD_DCN, 1>(p, st);
}
