// tc_common.cuh -- sm_100a tensor-core plumbing: tcgen05.mma / TMEM / mbarrier wrappers (inline PTX) and the
// shared-memory operand layout used by the tensor-core MLP paths.
//
// Precision scheme ("fp16x3"): an fp32 operand x is split as x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// (22 mantissa bits in total).  A product a*w is evaluated as  a_hi*w_hi + a_hi*w_lo + a_lo*w_hi  by kind::f16
// MMAs accumulating in fp32 TMEM; the dropped a_lo*w_lo term is 2^-22 relative.  Weights are pre-scaled by 2^8 so
// their lo halves stay normal in fp16.  Measured: MLP outputs within ~2e-6 of an fp64 evaluation relative to the
// output scale (the path's parity bar is 1e-5).  A bf16 variant of the helpers is kept for reference: its 16
// mantissa bits gave ~1e-5 and failed the bar for unit-variance emissions.
//
// Operand layout (both A = activations [rows x K] and B = weights [N x K], K-major, no swizzle):
//   element (r, k) lives at  base + (k/8) * (ROWS*16 B) + r * 16 B + (k%8) * 2 B
// i.e. K-groups of 8 elements are the slow index; inside a K-group every row owns one 16-byte chunk.  In UMMA
// terms the 8x16B core matrices of consecutive row groups are contiguous (SBO = 128 B) and consecutive K-groups
// are ROWS*16 B apart (LBO).  A thread that owns row r writes 8 consecutive k of its row with ONE conflict-free
// 16-byte store.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace psvo {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (K-major, SWIZZLE_NONE) --------------------------------------------
// bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ---- instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major -----------------------
// [4,6) c_format=1(f32) | [7,10) a_format=1(bf16) | [10,13) b_format=1(bf16) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// same, fp16 x fp16 -> fp32  (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// same with the A operand read from tensor memory (M = 128: row r = lane r, two fp16 K-elements per 32-bit column,
// so one K = 16 instruction reads 8 columns starting at a_tmem)
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// all MMAs issued so far by this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
               :: "r"(smem_u32(mbar)) : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------------------
// one full warp allocates `cols` (power of two >= 32) columns; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// make this thread's st.shared visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// this warp's 32 lanes x 16 consecutive columns, one row (lane) per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// this warp's 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// this warp's 32 lanes x 32 consecutive columns written from registers (the store is complete after tmem_st_wait)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
         "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
         "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
         "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// this warp's 32 lanes x 16 / x 8 consecutive columns written from registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
         "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// split form for software pipelining: issue the load of 32 columns now, fence it later.  The wait takes the destination
// registers as in/out operands so that no consumer can be scheduled above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
        "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
        "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :: "memory");
}

// 16-column variant of the split load
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
      :: "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------------------
// one lane of a fully converged warp; unlike `lane == 0` ptxas knows the guarded region runs on a single thread and
// emits the tcgen05 instructions without a per-active-lane serialisation loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(mbar)) : "memory");
}
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(uint64_t* mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}
// pure polling wait (no hardware sleep): lowest wake-up latency, costs issue slots while waiting
__device__ __forceinline__ void mbar_spin(uint64_t* mbar, uint32_t parity) {
  while (!mbar_test(mbar, parity)) {}
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}\n"
      :: "r"(smem_u32(mbar)), "r"(parity) : "memory");
}

// producer side of a transaction barrier: one arrival + `bytes` expected from bulk copies
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D, size multiple of 16, 16-byte aligned), completion counted on `mbar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

// ---- bf16 hi/lo split ---------------------------------------------------------------------------------------
// 8 consecutive k of one row -> two 16-byte chunks (hi, lo)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp16 variant of split8 (11 mantissa bits per half: hi + lo carries 22 bits).  |x| must stay below 65504.
__device__ __forceinline__ void split8_f16(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Truncating fp16 split for the hot epilogue: hi = x with the mantissa cut to fp16's 10 stored bits (one LOP3,
// exactly representable in fp16 for |x| in fp16's normal range), lo = fp16(x - hi) (exact subtraction, one
// rounding).  Residual <= 2^-21 |x|; no fp16 -> fp32 conversion is needed to form lo.
__device__ __forceinline__ void split8_f16_trunc(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float h0 = __uint_as_float(__float_as_uint(x[2 * i]) & 0xFFFFE000u);
    const float h1 = __uint_as_float(__float_as_uint(x[2 * i + 1]) & 0xFFFFE000u);
    const __half2 hh = __floats2half2_rn(h0, h1);
    const __half2 ll = __floats2half2_rn(x[2 * i] - h0, x[2 * i + 1] - h1);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ReLU + truncating fp16 split of 8 raw pre-activations in six instructions per pair instead of eight, spread over the
// FMA and ALU pipes.  cvt.*.relu.f16x2 applies the ReLU inside the packing conversion, so no FMNMX is needed:
//   even pairs:  hi_f = x & 0xFFFFE000 (LOP3, ALU) ; lo_f = x - hi_f ; hi = pack.relu(hi_f) ; lo = pack.relu(lo_f)
//   odd pairs :  hi = pack.rz.relu(x) ; hi_f = float(hi) (HADD2.F32, FMA pipe) ; lo_f = x - hi_f ; lo = pack.relu(lo_f)
// For x < 0 both give hi = lo = 0 (hi_f and lo_f are <= 0); for x >= 0 both give hi = trunc11(x), lo = rn(x - hi).
__device__ __forceinline__ uint32_t pack_relu_rn(float x0, float x1) {
  uint32_t r; asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0)); return r;
}
__device__ __forceinline__ uint32_t pack_relu_rz(float x0, float x1) {
  uint32_t r; asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0)); return r;
}
__device__ __forceinline__ void split8_relu_f16(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x0 = x[2 * i], x1 = x[2 * i + 1];
    if (i & 1) {
      h[i] = pack_relu_rz(x0, x1);
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
      l[i] = pack_relu_rn(x0 - hf.x, x1 - hf.y);
    } else {
      const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
      const float h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
      h[i] = pack_relu_rn(h0, h1);
      l[i] = pack_relu_rn(x0 - h0, x1 - h1);
    }
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp16 weight staging: W * scale split as hi + lo (scale is a power of two, undone in the epilogue)
__device__ __forceinline__ void stage_weight_f16x2(const float* __restrict__ W, int n_in, int n_out, int KP, int NP, float scale,
                                                   unsigned char* hi, unsigned char* lo) {
  for (int i = threadIdx.x; i < NP * KP; i += blockDim.x) {
    const int n = i % NP, k = i / NP;
    const float w = (n < n_out && k < n_in) ? __ldg(W + (size_t)k * n_out + n) * scale : 0.0f;
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const size_t off = (size_t)(k >> 3) * (NP * 16) + (size_t)n * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(hi + off) = h;
    *reinterpret_cast<__half*>(lo + off) = l;
  }
}

// same, hi and lo stacked along N inside one operand: per K-group NP "hi" rows followed by NP "lo" rows, so that
// one MMA with N = 2*NP evaluates a*w_hi and a*w_lo side by side, and an MMA with N = NP on the same descriptor
// (same LBO) sees only the hi rows.
__device__ __forceinline__ void stage_weight_f16x2_stacked(const float* __restrict__ W, int n_in, int n_out, int KP, int NP,
                                                           float scale, unsigned char* dst) {
  for (int i = threadIdx.x; i < NP * KP; i += blockDim.x) {
    const int n = i % NP, k = i / NP;
    const float w = (n < n_out && k < n_in) ? __ldg(W + (size_t)k * n_out + n) * scale : 0.0f;
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const size_t off = (size_t)(k >> 3) * (2 * NP * 16) + (size_t)n * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(dst + off) = h;
    *reinterpret_cast<__half*>(dst + off + NP * 16) = l;
  }
}

// stage an fp32 Keras kernel W[in][out] (row-major, `out` contiguous) as the B operand [N = out rows][K = in],
// hi and lo, zero padded to (NP rows, KP k).  Cooperative over the block.
__device__ __forceinline__ void stage_weight_bf16x2(const float* __restrict__ W, int n_in, int n_out, int KP, int NP,
                                                    unsigned char* hi, unsigned char* lo) {
  for (int i = threadIdx.x; i < NP * KP; i += blockDim.x) {
    const int n = i % NP, k = i / NP;                       // consecutive threads -> consecutive out (coalesced read)
    const float w = (n < n_out && k < n_in) ? __ldg(W + (size_t)k * n_out + n) : 0.0f;
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
    const size_t off = (size_t)(k >> 3) * (NP * 16) + (size_t)n * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo + off) = l;
  }
}

}  // namespace tc
}  // namespace psvo
