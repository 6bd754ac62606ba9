// hc_tracker_tp.cuh — the tracker with TWO PATHS PER WARP and TWO MATRIX ROWS PER LANE (included by hc_tracker.cu, inside its
// anonymous namespace; same arithmetic helpers, same arithmetic spec, results bit-identical to the one-path-per-warp kernel).
//
// Why: the one-path kernel is bounded by the shared-memory data pipe (profiles/ncu_r2.md: 84 % of its wavefront rate), and a
// third of those wavefronts are pivot rows that every lane loads to update ONE matrix row.  Here a half-warp (16 lanes) tracks a
// path, lane l of the half holds two rows ("row slots" 0 and 1, generator: build_tp) and the two variables l and l + 16, so a
// pivot row is loaded once per lane and applied to two rows, and one warp instruction serves two paths.
//
// Control: the two halves of a warp are two independent state machines (FETCH a path -> STEP prologue -> [STAGE -> POST]* ->
// finish), advanced in divergent per-half code with half-warp masks; the expensive STAGE body (table rebuild, x-products,
// Jacobian / right-hand side, 30x30 solve) is executed by the whole warp, each half on its own path's data.
#include "hc_problem_gen_tp.h"

#ifndef HC_TP_LOCKSTEP
#define HC_TP_LOCKSTEP 1
#endif
#ifndef HC_TP_WARPS
#define HC_TP_WARPS 16          // one persistent CTA per SM
#endif
constexpr int TP_WARPS = HC_TP_WARPS;
constexpr int TP_THREADS = TP_WARPS * 32;
constexpr int TP_NBUF = HCT_MAX_BUF;

__device__ const uint32_t g_tbl_tp[HCT_TBL_ROWS * 16] = HCT_TBL_INIT;
__device__ const uint32_t g_laneinfo_tp[32] = HCT_LANEINFO_INIT;      // [lane][row slot]: nz mask over register slots | selector bits << 20
__device__ const uint32_t g_lanecols_tp[32] = HCT_LANECOLS_INIT;      // [lane][row slot]: private column met on every level, 5 bits each
__device__ const uint32_t g_lanelvl_tp[32] = HCT_LANELVL_INIT;        // [lane][2]: 16 bits per level: active0 | active1<<1 | merged<<2 | buf0<<3 | buf1<<6
__device__ const int g_row_at_tp[32] = HCT_ROW_AT_INIT;              // [lane][row slot]

struct __align__(16) PathSmem {            // everything one path keeps on chip besides its two matrix rows per lane
  float2 p[36];                            // parameter homotopy p(t); p[33] == 1
  float2 dp[36];                           // target - start
  float2 xp[HCG_XP_TOTAL];                 // [0,32): evaluation point (x[30] == 1, x[31] == 0), then pair and triple products
  float2 cq[96];
  float2 dq[96];
  float2 row[2][TP_NBUF][ROWBUF_C];        // pivot-row broadcast, double buffered, one buffer per pivot group of a level
  float2 delta[32];                        // solution of the linear system, natural variable order
  float2 xl[32];                           // last accepted point, by variable
  float2 xs[32];                           // RK4 accumulator, by variable
};
constexpr size_t TP_SMEM_TBL = (size_t)HCT_TBL_ROWS * 16 * sizeof(uint32_t);
constexpr size_t TP_SMEM_SP = 36 * sizeof(float2);
constexpr size_t TP_SMEM_BYTES = TP_SMEM_TBL + TP_SMEM_SP + (size_t)TP_WARPS * 2 * sizeof(PathSmem);
static_assert(TP_SMEM_TBL % 16 == 0 && sizeof(PathSmem) % 16 == 0, "16-byte aligned carve-up");
static_assert(TP_SMEM_BYTES <= 227 * 1024, "one CTA per SM must fit the 227 KB of shared memory");

// ---- table builders (16 lanes per path; `on` = this half rebuilds) -------------------------------------------------------
__device__ __forceinline__ void tp_param_homotopy(PathSmem& w, const float2* __restrict__ s_sp, const float2* __restrict__ tgt, const int l,
                                                  const float t, const bool on)
{
  const float omt = __fsub_rn(1.0f, t);
  if (on) {
    w.p[l] = c_fma_s(tgt[l], t, c_scale(omt, s_sp[l]));
    w.p[l + 16] = c_fma_s(tgt[l + 16], t, c_scale(omt, s_sp[l + 16]));
    if (l == 0) w.p[32] = c_fma_s(tgt[32], t, c_scale(omt, s_sp[32]));
  }
}
__device__ __forceinline__ void tp_build_cq(PathSmem& w, const uint32_t* __restrict__ s_tbl, const int l, const bool on)
{
#pragma unroll
  for (int r = 0; r < HCT_CQ_ROUNDS; r++) {
    const uint32_t wd = s_tbl[(HCT_TBL_CQ + r) * 16 + l];
    const float2 pa = lds_off<float2>(w.p, wd & 0x1ffu);
    const float2 pb = lds_off<float2>(w.p, (wd >> 9) & 0x1ffu);
    const float coef = (float)((int)((wd >> 18) & 7u) - 2);
    const float2 v = c_mul(pa, pb);
    if (on && r * 16 + l < HCG_NUM_CQ) w.cq[r * 16 + l] = c_scale(coef, v);
  }
}
__device__ __forceinline__ void tp_build_dq(PathSmem& w, const uint32_t* __restrict__ s_tbl, const int l, const bool on)
{
#pragma unroll
  for (int r = 0; r < HCT_DQ_ROUNDS; r++) {
    const uint32_t wd = s_tbl[(HCT_TBL_DQ + r) * 16 + l];
    const unsigned oa = wd & 0x1ffu, ob = (wd >> 9) & 0x1ffu;
    const float2 pa = lds_off<float2>(w.p, oa), pb = lds_off<float2>(w.p, ob);
    const float2 da = lds_off<float2>(w.dp, oa), db = lds_off<float2>(w.dp, ob);
    const float coef = (float)((int)((wd >> 18) & 7u) - 2);
    const float2 v = c_add(c_mul(da, pb), c_mul(db, pa));
    if (on && r * 16 + l < HCG_NUM_DQ) w.dq[r * 16 + l] = c_scale(-coef, v);
  }
}
__device__ __forceinline__ void tp_build_xp(PathSmem& w, const uint32_t* __restrict__ s_tbl, const int l)
{
#pragma unroll
  for (int r = 0; r < HCT_PAIR_ROUNDS; r++) {
    const uint32_t wd = s_tbl[(HCT_TBL_PAIR + r) * 16 + l];
    const float2 v = c_mul(lds_off<float2>(w.xp, wd & 0xffu), lds_off<float2>(w.xp, (wd >> 8) & 0xffu));
    if (r * 16 + l < HCG_XP_TRI0 - HCG_XP_PAIR0) w.xp[HCG_XP_PAIR0 + r * 16 + l] = v;
  }
#pragma unroll
  for (int r = 0; r < HCT_TRI_ROUNDS; r++) {
    const uint32_t wd = s_tbl[(HCT_TBL_TRI + r) * 16 + l];
    float2 v = c_mul(lds_off<float2>(w.xp, wd & 0xffu), lds_off<float2>(w.xp, (wd >> 8) & 0xffu));
    v = c_mul(v, lds_off<float2>(w.xp, (wd >> 16) & 0xffu));
    if (r * 16 + l < HCG_XP_TOTAL - HCG_XP_TRI0) w.xp[HCG_XP_TRI0 + r * 16 + l] = v;
  }
}

// ---- evaluators: one row slot at a time (R = 0, 1), same term order per matrix entry as the one-path kernel ----------------
template <int R>
__device__ __forceinline__ void tp_eval_Hx(float2 (&A)[NSLOT], const PathSmem& w, const uint32_t* __restrict__ s_tbl, const int l, const uint32_t laneinfo)
{
  constexpr int kBeg[2 * HCG_NUM_CLASSES + 1] = HCT_HX_BEGIN_INIT;
  float2 acc[HCG_NUM_CLASSES];
#pragma unroll
  for (int c = 0; c < HCG_NUM_CLASSES; c++) {
    float2 a = make_float2(0.0f, 0.0f);
    const uint32_t* rows = s_tbl + kBeg[2 * c + R] * 16 + l;
#pragma unroll kHxUnroll
    for (int s = 0; s < kBeg[2 * c + R + 1] - kBeg[2 * c + R]; s++) {
      const uint32_t wd = rows[s * 16];
      a = c_add(a, c_mul(lds_off<float2>(w.cq, wd & 0xffffu), lds_off<float2>(w.xp, wd >> 16)));
    }
    acc[c] = a;
  }
#define HC_TP_SLOT_OUT(T, CA, CB, SB)                                                                    \
  {                                                                                                      \
    float2 v = acc[CA];                                                                                  \
    if ((CB) >= 0) { if ((laneinfo >> (20 + ((SB) < 0 ? 0 : (SB)))) & 1u) v = acc[(CB) < 0 ? 0 : (CB)]; } \
    A[T] = ((laneinfo >> (T)) & 1u) ? v : make_float2(0.0f, 0.0f);                                       \
  }
  HCT_HX_SCATTER_LIST(HC_TP_SLOT_OUT)
#undef HC_TP_SLOT_OUT
}
__device__ __forceinline__ float2 tp_eval_rhs(const PathSmem& w, const float2* __restrict__ coef_tbl, const uint32_t* __restrict__ s_rows, const int l)
{
  float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll 4
  for (int s = 0; s < HCT_RHS_SLOTS; s++) {
    const uint32_t wd = s_rows[s * 16 + l];
    acc = c_add(acc, c_mul(lds_off<float2>(coef_tbl, wd & 0xffffu), lds_off<float2>(w.xp, wd >> 16)));
  }
  return acc;
}

// ---- solve -----------------------------------------------------------------------------------------------------------------
struct TpState {
  uint32_t alive0, alive1;     // all ones while the row is a pivot candidate
  uint32_t min_mx;
  int mystep0, mystep1;
  float2 rsave0, rsave1;
  uint32_t ck0, ck1, mx0, mx1; // look-ahead of the next pivot search, per row slot
  float2 r0, r1;               // 1 / (my entry in the next pivot column), per row slot
};

// maximum over the three lanes of my segment (a row that is no candidate contributes 0)
__device__ __forceinline__ uint32_t tp_seg_max(const uint32_t v, const int segbase)
{
  uint32_t m = __shfl_sync(FULL, v, segbase);
  m = max(m, __shfl_sync(FULL, v, segbase + 1));
  return max(m, __shfl_sync(FULL, v, segbase + 2));
}
__device__ __forceinline__ uint32_t tp_half_max(const uint32_t v)      // maximum over the 16 lanes of my half-warp
{
  uint32_t m = v;
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(FULL, m, off));
  return m;
}

// level word of level T: active0 | active1<<1 | merged<<2 | buf0<<3 | buf1<<6
template <int T> __device__ __forceinline__ uint32_t tp_lvl(const uint32_t lv01, const uint32_t lv23)
{ return ((T < 2 ? lv01 : lv23) >> (16 * (T & 1))) & 0xffffu; }

// look-ahead for slot TN: candidate keys of both row slots, the group maxima and both reciprocals
template <int TN>
__device__ __forceinline__ void tp_lookahead(const float2 (&A0)[NSLOT], const float2 (&A1)[NSLOT], TpState& st, const int segbase,
                                             const uint32_t lv01, const uint32_t lv23, const uint32_t rowkey0, const uint32_t rowkey1)
{
  if constexpr (TN < NSP) {
    const uint32_t lw = tp_lvl<TN>(lv01, lv23);
    st.ck0 = cand_key(A0[TN], (lw & 1u) ? st.alive0 : 0u, rowkey0);
    st.ck1 = cand_key(A1[TN], (lw & 2u) ? st.alive1 : 0u, rowkey1);
    const uint32_t mA = tp_seg_max(st.ck0, segbase), mB = tp_seg_max(st.ck1, segbase);
    const uint32_t mm = max(mA, mB);
    st.mx0 = (lw & 4u) ? mm : mA;
    st.mx1 = (lw & 4u) ? mm : mB;
  } else {
    st.ck0 = cand_key(A0[TN], st.alive0, rowkey0);
    st.ck1 = cand_key(A1[TN], st.alive1, rowkey1);
    st.mx0 = st.mx1 = tp_half_max(max(st.ck0, st.ck1));
  }
  st.r0 = c_recip(A0[TN]);
  st.r1 = c_recip(A1[TN]);
}

template <int T0>
__device__ __forceinline__ void tp_store_row(const float2 (&A)[NSLOT], const float2 b, const float2 r, const uint32_t rb, const uint32_t p)
{
#pragma unroll
  for (int t = ((T0 + 1) & ~1); t < NSLOT; t += 2) {
    if (!(touches<T0>(t) || touches<T0>(t + 1))) continue;
    sts_v4_p(rb + t * 8, A[t].x, A[t].y, A[t + 1].x, A[t + 1].y, p);
  }
  sts_v4_p(rb + RB_RHS * 8, b.x, b.y, r.x, r.y, p);
}

// One elimination step on slot T for both row slots of the lane.  SPLIT (level 0 of this problem): the two row slots may belong to
// different pivot groups, so each slot loads its own group's pivot row; otherwise both rows share one pivot row and one load.
template <int T, bool EXACT>
__device__ __forceinline__ void tp_solve_step(float2 (&A0)[NSLOT], float2 (&A1)[NSLOT], float2& b0, float2& b1, TpState& st, const uint32_t row_sa,
                                              const int segbase, const uint32_t lv01, const uint32_t lv23, const uint32_t rowkey0, const uint32_t rowkey1)
{
  constexpr bool SEGMENTED = (T < NSP);
  const uint32_t lw = SEGMENTED ? tp_lvl<T < NSP ? T : 0>(lv01, lv23) : 0x7u;       // warp-wide steps: both active, one group, buffer 0
  const bool act0 = (lw & 1u) != 0u, act1 = (lw & 2u) != 0u;
  const bool piv0 = act0 && (st.ck0 == st.mx0) && (st.ck0 != 0u);
  const bool piv1 = act1 && (st.ck1 == st.mx1) && (st.ck1 != 0u);
  st.min_mx = min(st.min_mx, act0 ? st.mx0 : 0xffffffffu);
  st.min_mx = min(st.min_mx, act1 ? st.mx1 : 0xffffffffu);
  const uint32_t rbp = row_sa + (uint32_t)(T & 1) * (TP_NBUF * ROWBUF_BYTES);
  const uint32_t rb0 = rbp + ((lw >> 3) & 7u) * ROWBUF_BYTES, rb1 = rbp + ((lw >> 6) & 7u) * ROWBUF_BYTES;
  st.alive0 = piv0 ? 0u : st.alive0;
  st.alive1 = piv1 ? 0u : st.alive1;
  st.rsave0.x = piv0 ? st.r0.x : st.rsave0.x; st.rsave0.y = piv0 ? st.r0.y : st.rsave0.y;
  st.rsave1.x = piv1 ? st.r1.x : st.rsave1.x; st.rsave1.y = piv1 ? st.r1.y : st.rsave1.y;
  st.mystep0 = piv0 ? T : st.mystep0;
  st.mystep1 = piv1 ? T : st.mystep1;
  tp_store_row<T>(A0, b0, st.r0, rb0, (uint32_t)piv0);
  tp_store_row<T>(A1, b1, st.r1, rb1, (uint32_t)piv1);
  __syncwarp();
  const bool upd0 = act0 && !piv0 && (A0[T].x != 0.0f || A0[T].y != 0.0f);
  const bool upd1 = act1 && !piv1 && (A1[T].x != 0.0f || A1[T].y != 0.0f);
  constexpr bool SPLIT = SEGMENTED && (T == 0);          // generator: only level 0 has segments whose two row slots are separate groups
  const float4 br0 = lds_v4(rb0 + RB_RHS * 8);
  const float4 br1 = SPLIT ? lds_v4(rb1 + RB_RHS * 8) : br0;
  const MultX<EXACT> m0 = make_multx<EXACT>(c_mul(A0[T], make_float2(br0.z, br0.w)), upd0);
  const MultX<EXACT> m1 = make_multx<EXACT>(c_mul(A1[T], make_float2(br1.z, br1.w)), upd1);
  // slot T+1 first, then the look-ahead for it, then the rest of the row
  if constexpr (T + 1 < NSLOT) {
    if constexpr (touches<T>(T + 1)) {
      const float2 u0 = lds_v2(rb0 + (T + 1) * 8);
      const float2 u1 = SPLIT ? lds_v2(rb1 + (T + 1) * 8) : u0;
      c_msub_x<EXACT>(A0[T + 1], m0, u0);
      c_msub_x<EXACT>(A1[T + 1], m1, u1);
    }
    tp_lookahead<T + 1>(A0, A1, st, segbase, lv01, lv23, rowkey0, rowkey1);
  }
  // slot T+2 alone when it is the odd half of a pair, then aligned pairs (one 128-bit load each)
  constexpr int TA = ((T + 2) & 1) ? T + 3 : T + 2;       // first even slot >= T + 2
  if constexpr (TA != T + 2 && T + 2 < NSLOT) {
    if constexpr (touches<T>(T + 2)) {
      const float2 u0 = lds_v2(rb0 + (T + 2) * 8);
      const float2 u1 = SPLIT ? lds_v2(rb1 + (T + 2) * 8) : u0;
      c_msub_x<EXACT>(A0[T + 2], m0, u0);
      c_msub_x<EXACT>(A1[T + 2], m1, u1);
    }
  }
#pragma unroll
  for (int t = TA; t < NSLOT; t += 2) {
    if (!(touches<T>(t) || touches<T>(t + 1))) continue;
    const float4 u0 = lds_v4(rb0 + t * 8);
    const float4 u1 = SPLIT ? lds_v4(rb1 + t * 8) : u0;
    if (touches<T>(t)) { c_msub_x<EXACT>(A0[t], m0, make_float2(u0.x, u0.y)); c_msub_x<EXACT>(A1[t], m1, make_float2(u1.x, u1.y)); }
    if (touches<T>(t + 1)) { c_msub_x<EXACT>(A0[t + 1], m0, make_float2(u0.z, u0.w)); c_msub_x<EXACT>(A1[t + 1], m1, make_float2(u1.z, u1.w)); }
  }
  c_msub_x<EXACT>(b0, m0, make_float2(br0.x, br0.y));
  c_msub_x<EXACT>(b1, m1, make_float2(br1.x, br1.y));
}
template <int T, bool EXACT>
__device__ __forceinline__ void tp_solve_steps(float2 (&A0)[NSLOT], float2 (&A1)[NSLOT], float2& b0, float2& b1, TpState& st, const uint32_t row_sa,
                                               const int segbase, const uint32_t lv01, const uint32_t lv23, const uint32_t rowkey0, const uint32_t rowkey1)
{
  if constexpr (T < NSLOT) {
    tp_solve_step<T, EXACT>(A0, A1, b0, b1, st, row_sa, segbase, lv01, lv23, rowkey0, rowkey1);
    tp_solve_steps<T + 1, EXACT>(A0, A1, b0, b1, st, row_sa, segbase, lv01, lv23, rowkey0, rowkey1);
  }
}

// One linear stage for both halves of the warp: evaluate, solve, leave the solution in w.delta (natural variable order).
template <bool EXACT>
__device__ __forceinline__ void tp_stage(PathSmem& w, const uint32_t* __restrict__ s_tbl, const bool pred, const uint32_t row_sa, const int l,
                                         const int lane, const uint32_t info0, const uint32_t info1, const uint32_t cols0, const uint32_t cols1,
                                         const uint32_t lv01, const uint32_t lv23, const uint32_t rowkey0, const uint32_t rowkey1, const bool has0, const bool has1)
{
  constexpr int kH[2] = HCT_TBL_H_INIT, kHt[2] = HCT_TBL_HT_INIT;
  float2 A0[NSLOT], A1[NSLOT];
  tp_eval_Hx<0>(A0, w, s_tbl, l, info0);
  tp_eval_Hx<1>(A1, w, s_tbl, l, info1);
  const float2* coef = pred ? w.dq : w.cq;
  float2 b0 = tp_eval_rhs(w, coef, s_tbl + (pred ? kHt[0] : kH[0]) * 16, l);
  float2 b1 = tp_eval_rhs(w, coef, s_tbl + (pred ? kHt[1] : kH[1]) * 16, l);
  if (!has0) b0 = make_float2(0.0f, 0.0f);
  if (!has1) b1 = make_float2(0.0f, 0.0f);
  const int segbase = (lane & 16) + min(l / 3, HCG_NSEG - 1) * 3;      // first lane of my 3-lane segment (lane 15 rides with the last one)
  TpState st;
  st.alive0 = has0 ? 0xffffffffu : 0u; st.alive1 = has1 ? 0xffffffffu : 0u;
  st.min_mx = 0xffffffffu; st.mystep0 = st.mystep1 = 0;
  st.rsave0 = st.rsave1 = make_float2(0.0f, 0.0f);
  tp_lookahead<0>(A0, A1, st, segbase, lv01, lv23, rowkey0, rowkey1);
  tp_solve_steps<0, EXACT>(A0, A1, b0, b1, st, row_sa, segbase, lv01, lv23, rowkey0, rowkey1);
  // a pivot column that is exactly zero in my path makes its whole solution NaN (arithmetic spec)
  const uint32_t sing = __ballot_sync(FULL, st.min_mx < 32u);
  const bool any_singular = ((sing >> (lane & 16)) & 0xffffu) != 0u;
  const float qnan = __int_as_float(0x7fffffff);
  if (has0) {
    const int col = (st.mystep0 < NSP) ? (int)((cols0 >> (5 * st.mystep0)) & 31u) : HCG_K1 + (st.mystep0 - NSP);
    w.delta[col] = any_singular ? make_float2(qnan, qnan) : c_mul(b0, st.rsave0);
  }
  if (has1) {
    const int col = (st.mystep1 < NSP) ? (int)((cols1 >> (5 * st.mystep1)) & 31u) : HCG_K1 + (st.mystep1 - NSP);
    w.delta[col] = any_singular ? make_float2(qnan, qnan) : c_mul(b1, st.rsave1);
  }
  __syncwarp();
}
// the rare exact repeat lives in its own never-inlined instance (cold code)
__device__ __noinline__ void tp_stage_exact(PathSmem& w, const uint32_t* __restrict__ s_tbl, const bool pred, const uint32_t row_sa, const int l,
                                            const int lane, const uint32_t info0, const uint32_t info1, const uint32_t cols0, const uint32_t cols1,
                                            const uint32_t lv01, const uint32_t lv23, const uint32_t rowkey0, const uint32_t rowkey1, const bool has0, const bool has1)
{ tp_stage<true>(w, s_tbl, pred, row_sa, l, lane, info0, info1, cols0, cols1, lv01, lv23, rowkey0, rowkey1, has0, has1); }

// sum over the 30 variables in the order of the spec's xor butterfly: level 16 is the lane-local add of (l, l + 16)
__device__ __forceinline__ float tp_half_sum(const float v0, const float v1, const unsigned hmask)
{
  float v = __fadd_rn(v0, v1);
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(hmask, v, off));
  return v;
}

// ---------------------------------------------------------------------------------------------------------------------------
enum { TP_FETCH = 0, TP_STEP = 1, TP_STAGE = 2, TP_POST = 3, TP_DONE = 4 };

__global__ void __launch_bounds__(TP_THREADS, 1) hc_track_tp_kernel(const Params P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_tbl = reinterpret_cast<uint32_t*>(smem_raw);
  float2* s_sp = reinterpret_cast<float2*>(smem_raw + TP_SMEM_TBL);
  PathSmem* s_path = reinterpret_cast<PathSmem*>(smem_raw + TP_SMEM_TBL + TP_SMEM_SP);

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int half = lane >> 4, l = lane & 15;
  const unsigned hmask = 0xffffu << (16 * half);
  for (int i = threadIdx.x; i < HCT_TBL_ROWS * 16; i += TP_THREADS) s_tbl[i] = g_tbl_tp[i];
  if (threadIdx.x < NP1) s_sp[threadIdx.x] = P.start_params[threadIdx.x];
  PathSmem& w = s_path[2 * wid + half];
  const uint32_t row_sa = (uint32_t)__cvta_generic_to_shared(&w.row[0][0][0]);
  const uint32_t info0 = g_laneinfo_tp[2 * l], info1 = g_laneinfo_tp[2 * l + 1];
  const uint32_t cols0 = g_lanecols_tp[2 * l], cols1 = g_lanecols_tp[2 * l + 1];
  const uint32_t lv01 = g_lanelvl_tp[2 * l], lv23 = g_lanelvl_tp[2 * l + 1];
  const int row0 = g_row_at_tp[2 * l], row1 = g_row_at_tp[2 * l + 1];
  const bool has0 = row0 >= 0, has1 = row1 >= 0;
  const uint32_t rowkey0 = (uint32_t)(31 - row0) & 31u, rowkey1 = (uint32_t)(31 - row1) & 31u;
  const bool var1 = (l + 16) < N;                        // my second variable exists (lanes 14, 15 hold the constants x[30], x[31])
  for (int i = l; i < (int)(sizeof(w.row) / sizeof(float2)); i += 16) (&w.row[0][0][0])[i] = make_float2(0.0f, 0.0f);
  if (l == 0) { w.xp[N] = make_float2(1.0f, 0.0f); w.xp[N + 1] = make_float2(0.0f, 0.0f); w.p[NP1 - 1] = make_float2(1.0f, 0.0f); }
  if (l < 4) { w.p[32 + l] = (l == 1) ? make_float2(1.0f, 0.0f) : make_float2(0.0f, 0.0f); w.dp[32 + l] = make_float2(0.0f, 0.0f); }
  for (int i = l; i < 96; i += 16) { w.cq[i] = make_float2(0.0f, 0.0f); w.dq[i] = make_float2(0.0f, 0.0f); }
  for (int i = l; i < HCG_XP_TOTAL; i += 16) if (i != N) w.xp[i] = make_float2(0.0f, 0.0f);
  __syncthreads();

  const float c6_1 = (float)(1.0 / 6.0), c6_2 = (float)(2.0 / 6.0);
  const bool prune = (P.flags & HCB200_FLAG_PRUNE_PATHS) != 0u;

  // per-half state (identical in the 16 lanes of a half)
  int phase = TP_FETCH, path = 0, hyp = 0, step = 0, st = 0, counter = 0, reason = 3;
  uint32_t cnt = 0u;
  float t0 = 0.0f, t_step = 0.0f, delta_t = 0.01f, half_dt = 0.0f;
  bool end_zone = false, ok = false, inf_fail = false, check_depths = true, rebuild = false;
  float2 xt0 = make_float2(0.0f, 0.0f), xt1 = make_float2(0.0f, 0.0f);

  for (;;) {
    // ---- per-half control: advance until this half has a stage to run (or is out of work) ------------------------------------
    while (phase != TP_STAGE && phase != TP_DONE) {
      if (phase == TP_FETCH) {
        int pth = 0;
        if (l == 0) pth = (int)atomicAdd(&P.ws->next_path, 1u);
        pth = __shfl_sync(hmask, pth, 16 * half);
        if (pth >= P.n_paths) { phase = TP_DONE; break; }
        path = pth; hyp = path / TRACKS;
        const int sidx = path - hyp * TRACKS;
        xt0 = P.start_sols[sidx * (N + 1) + l];
        xt1 = var1 ? P.start_sols[sidx * (N + 1) + l + 16] : make_float2(0.0f, 0.0f);
        w.dp[l] = P.diff_params[hyp * NP1 + l];
        w.dp[l + 16] = P.diff_params[hyp * NP1 + l + 16];
        if (l < 2) w.dp[32 + l] = P.diff_params[hyp * NP1 + 32 + l];
        w.xl[l] = xt0; w.xl[l + 16] = xt1; w.xs[l] = xt0; w.xs[l + 16] = xt1;
        t0 = 0.0f; delta_t = 0.01f; end_zone = false; ok = false; inf_fail = false; check_depths = true;
        counter = 0; cnt = 0u; reason = 3; step = 0;
        phase = TP_STEP;
      } else if (phase == TP_STEP) {                     // loop head of …TrunPaths.cu:137-165
        bool finish = step > P.max_steps;
        if (!finish && !((double)t0 < 1.0 && (1.0 - (double)t0 > 0.0000001))) { reason = 0; finish = true; }
        if (!finish) {
          if (!end_zone && (double)fabsf(__fsub_rn(1.0f, t0)) <= 0.0500001) end_zone = true;
          if (prune) {
            if (check_depths) {
              const bool all_pos = (__ballot_sync(hmask, (l >= 8) || (xt0.x > 0.0f)) == hmask);
              if (t0 > 0.0f) check_depths = !all_pos;
            }
            if ((double)t0 > 0.95 && check_depths) { reason = 2; finish = true; }
          }
        }
        if (!finish) {
          if (end_zone) { const float r = fabsf(__fsub_rn(1.0f, t0)); if (delta_t > r) delta_t = r; }
          else { const double r = fabs(0.95 - (double)t0); if ((double)delta_t > r) delta_t = (float)r; }
          t_step = t0;
          half_dt = __fmul_rn(0.5f, delta_t);
          cnt += 1u;
          st = 0;
          rebuild = !ok;                                 // `ok` still holds the verdict of the previous step (false at the start)
          phase = TP_STAGE;
        } else {                                         // the path is over: write it out (…TrunPaths.cu:282-286)
          const bool conv = ((double)t0 >= 1.0 || (1.0 - (double)t0 <= 0.0000001));
          if (conv && reason == 3) reason = 0;
          P.tracks[(size_t)path * (N + 1) + l] = xt0;
          if (var1) P.tracks[(size_t)path * (N + 1) + l + 16] = xt1;
          if (l == 14) P.tracks[(size_t)path * (N + 1) + N] = make_float2(1.0f, 0.0f);
          if (l == 0) {
            P.converged[path] = conv ? 1 : 0;
            P.infinity[path] = inf_fail ? 1 : 0;
            if (P.stats) {
              hcb200_path_stats s; s.steps = (int)(cnt & 1023u); s.pred_stages = 4 * s.steps; s.corr_stages = (int)((cnt >> 10) & 4095u);
              s.rejected_reason = (int)(cnt >> 22) | (reason << 16);
              P.stats[path] = s;
            }
          }
          phase = TP_FETCH;
        }
      } else {                                           // TP_POST: bookkeeping after stage `st` (…TrunPaths.cu:191-275)
        const float2 d0 = w.delta[l], d1 = w.delta[l + 16];
        bool step_end = false;
        if (st < 4) {
          const float2 s0 = w.xs[l], s1 = w.xs[l + 16];
          if (st < 3) {
            const float c6 = (st == 0) ? c6_1 : c6_2;
            const float sc = (st < 2) ? half_dt : delta_t;
            w.xs[l] = c_fma_s(c_scale(delta_t, d0), c6, s0);
            if (var1) w.xs[l + 16] = c_fma_s(c_scale(delta_t, d1), c6, s1);
            xt0 = c_fma_s(d0, sc, w.xl[l]);
            if (var1) xt1 = c_fma_s(d1, sc, w.xl[l + 16]);
            if (st != 1) t0 = __fadd_rn(t0, half_dt);
          } else {
            const float2 k0 = c_scale(delta_t, d0), k1 = c_scale(delta_t, d1);
            xt0 = c_add(s0, make_float2(__fdiv_rn(k0.x, 6.0f), __fdiv_rn(k0.y, 6.0f)));
            if (var1) xt1 = c_add(s1, make_float2(__fdiv_rn(k1.x, 6.0f), __fdiv_rn(k1.y, 6.0f)));
          }
          st++;
          rebuild = (st == 1 || st == 3);
          phase = TP_STAGE;
        } else {
          cnt += 1u << 10;
          xt0 = c_sub(xt0, d0);
          if (var1) xt1 = c_sub(xt1, d1);
          const float vd0 = __fmaf_rn(d0.x, d0.x, __fmul_rn(d0.y, d0.y));
          const float vd1 = var1 ? __fmaf_rn(d1.x, d1.x, __fmul_rn(d1.y, d1.y)) : 0.0f;
          const float vx0 = __fmaf_rn(xt0.x, xt0.x, __fmul_rn(xt0.y, xt0.y));
          const float vx1 = var1 ? __fmaf_rn(xt1.x, xt1.x, __fmul_rn(xt1.y, xt1.y)) : 0.0f;
          const float sum_d = tp_half_sum(vd0, vd1, hmask), sum_x = tp_half_sum(vx0, vx1, hmask);
          ok = (double)sum_d < 0.000001 * (double)sum_x;
          inf_fail = (double)sum_x > 1e14;
          if (inf_fail || ok || st + 1 >= 4 + P.max_corr) step_end = true;
          else { st++; rebuild = false; phase = TP_STAGE; }
        }
        if (step_end) {
          if (inf_fail) { reason = 1; step = P.max_steps + 1; }      // leave the step loop: TP_STEP writes the path out with reason 1
          else if (!ok) {                                // …TrunPaths.cu:257-275
            delta_t = __fmul_rn(delta_t, 0.5f);
            xt0 = w.xl[l]; xt1 = w.xl[l + 16];
            w.xs[l] = xt0; w.xs[l + 16] = xt1;
            counter = 0;
            t0 = t_step;
            cnt += 1u << 22;
            step++;
          } else {
            counter++;
            w.xl[l] = xt0; w.xl[l + 16] = xt1; w.xs[l] = xt0; w.xs[l + 16] = xt1;
            if (counter >= P.dt_inc) { counter = 0; delta_t = __fmul_rn(delta_t, 2.0f); }
            step++;
          }
          phase = TP_STEP;
        }
      }
    }
#if HC_TP_LOCKSTEP
    // all warps of the CTA enter the stage body together: they then walk the same instructions at the same time and share their
    // instruction-cache lines (the body is ~45 KB of straight-line code against a 32 KB L1.5 instruction cache)
    if (__syncthreads_and(phase == TP_DONE)) break;
#else
    __syncwarp();
    if (__all_sync(FULL, phase == TP_DONE)) break;
#endif
    const bool run = (phase == TP_STAGE);

    // ---- the stage body, both halves together ------------------------------------------------------------------------------------
    if (__any_sync(FULL, run && rebuild)) {
      const bool on = run && rebuild;
      tp_param_homotopy(w, s_sp, P.target_params + (size_t)hyp * NP1, l, t0, on);
      __syncwarp();
      tp_build_cq(w, s_tbl, l, on);
      tp_build_dq(w, s_tbl, l, on);
    }
    if (run) { w.xp[l] = xt0; if (var1) w.xp[l + 16] = xt1; }
    __syncwarp();
    tp_build_xp(w, s_tbl, l);
    __syncwarp();
    const bool pred = st < 4;
    tp_stage<false>(w, s_tbl, pred, row_sa, l, lane, info0, info1, cols0, cols1, lv01, lv23, rowkey0, rowkey1, has0, has1);
#if HC_PACKED
    {
      const float2 d0 = w.delta[l], d1 = w.delta[l + 16];
      const bool bad = run && (!(fabsf(d0.x) <= 3.402823466e+38f && fabsf(d0.y) <= 3.402823466e+38f) ||
                               (var1 && !(fabsf(d1.x) <= 3.402823466e+38f && fabsf(d1.y) <= 3.402823466e+38f)));
      if (__any_sync(FULL, bad)) {
        __syncwarp();
        tp_stage_exact(w, s_tbl, pred, row_sa, l, lane, info0, info1, cols0, cols1, lv01, lv23, rowkey0, rowkey1, has0, has1);
      }
    }
#endif
    if (run) phase = TP_POST;
  }
}
