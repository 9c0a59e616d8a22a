// rrt_packed.cuh -- render_kernel_p: the render kernel with TWO rays per thread in packed f32x2 registers, FMAD contract
// only (included by rrt_kernel.cuh when RRT_FMAD = 1).
//
// Why.  The scalar step loop is bound by register-file operand delivery, not by FLOPs or issue slots
// (profiles/r2_rf_model.md): an SMSP's register file has two banks (even / odd register index) that deliver one 32-bit
// operand per cycle each, so a scalar instruction costs max(1, even reads, odd reads) cycles -- a three-register FFMA is
// two cycles unless ptxas finds an operand to reuse, and a two-register FMUL is two cycles whenever its operands fall
// into the same bank.  The 201 instructions of one vacuum step cost 245-285 cycles depending on ptxas's register
// allocation luck.  Blackwell's packed FP32 instructions (FFMA2 / FMUL2 / FADD2, two IEEE binary32 operations per
// instruction on an aligned 64-bit register pair) take every operand as one even + one odd register, so their cost is
// max(2, distinct register pairs) cycles for two operations, whatever the allocation: 1.5 cycles per operation for a
// three-operand FFMA2, 1 otherwise -- and half the issue slots, which hides the loop's non-FP instructions.  Uniform
// operands (step sizes, the two RHS constants, 2.0, 0.5, 1.0) are broadcast from uniform registers or immediates and
// cost no vector-register read.  Each half is rounded separately with exactly the operations of the scalar code in
// rrt_device.cuh, so a ray's result does not depend on which kernel traced it (tests/test_gpu_packed.py: frames,
// planes and counters bit-identical to render_kernel's).
//
// Shape.  One warp = a 16x4-pixel tile; thread (lx, ly) owns pixels (2 lx, ly) and (2 lx + 1, ly).  The two rays step in
// lock-step (one iteration counter); a ray that ends is finalised at once (sky, effects, store) and its half is parked on
// an inert state -- at rest on the spin axis at r = 100, where the RHS is exactly zero -- until its partner ends.  The
// loop has the two phases of render_kernel: unchecked vacuum bursts while every ray of the warp is inside the burst
// window (constant step: all step-size operands are uniform), and checked iterations (per-ray zone logic, media
// samples, escape test; the step size is a per-half register pair there) otherwise.
#pragma once

namespace rrtp {
using rrt::Consts;
using rrt::V3;
using rrt::mk;

using rrt::f2::F2;   // .lo = ray A, .hi = ray B
using rrt::f2::pk;
using rrt::f2::upk;
using rrt::f2::bc;
using rrt::f2::neg2;
using rrt::f2::add2;
using rrt::f2::mul2;
using rrt::f2::fma2;
using rrt::f2::mul2_unfusable;
__device__ __forceinline__ float half_of(F2 a, int hf) { float l, h; upk(a, l, h); return hf ? h : l; }

struct V3x2 {
    F2 x, y, z;
};
__device__ __forceinline__ V3 half_of(const V3x2& a, int hf) { return mk(half_of(a.x, hf), half_of(a.y, hf), half_of(a.z, hf)); }
__device__ __forceinline__ void set_half(V3x2& a, int hf, V3 v) {
    float l, h;
    upk(a.x, l, h); a.x = hf ? pk(l, v.x) : pk(v.x, h);
    upk(a.y, l, h); a.y = hf ? pk(l, v.y) : pk(v.y, h);
    upk(a.z, l, h); a.z = hf ? pk(l, v.z) : pk(v.z, h);
}

// ---- the FMAD contract's primitives (include/rrt_device.cuh), two rays at a time ---------------------------------
// dot(a, b) = fma(a.z, b.z, fma(a.x, b.x, a.y * b.y))
__device__ __forceinline__ F2 dot3_2(const V3x2& a, const V3x2& b) { return fma2(a.z, b.z, fma2(a.x, b.x, mul2(a.y, b.y))); }
// |p|^2 of the loop header: (fma(y, y, x * x)) + z * z, the last product rounded on its own
__device__ __forceinline__ F2 norm2_loop_2(const V3x2& p) { return add2(fma2(p.y, p.y, mul2(p.x, p.x)), mul2_unfusable(p.z, p.z)); }
// x / y, correctly rounded: the branch-free fast path of div_rn_fast
__device__ __forceinline__ F2 div2(F2 x, F2 y) {
    float yl, yh;
    upk(y, yl, yh);
    const F2 ny = neg2(y);
    F2 r = pk(rrt::rcp_approx(yl), rrt::rcp_approx(yh));
    const F2 e = fma2(ny, r, bc(1.0f));
    r = fma2(r, e, r);
    const F2 q = mul2(x, r);
    const F2 rem = fma2(ny, q, x);
    return fma2(r, rem, q);
}
// sqrt(x), correctly rounded: sqrt_rn_fast
__device__ __forceinline__ F2 sqrt2(F2 x) {
    float xl, xh;
    upk(x, xl, xh);
    const F2 y = pk(rrt::rsqrt_approx(xl), rrt::rsqrt_approx(xh));
    const F2 g = mul2(x, y);
    const F2 hlf = mul2(y, bc(0.5f));
    const F2 e = fma2(neg2(g), g, x);
    return fma2(e, hlf, g);
}
__device__ __forceinline__ V3x2 axpy2(const V3x2& y, const V3x2& x, F2 a) {   // y + x * a
    V3x2 o;
    o.x = fma2(x.x, a, y.x); o.y = fma2(x.y, a, y.y); o.z = fma2(x.z, a, y.z);
    return o;
}
__device__ __forceinline__ F2 rk4_sum2(F2 k1, F2 k2, F2 k3, F2 k4) { return add2(k1, fma2(bc(2.0f), k2, fma2(bc(2.0f), k3, k4))); }

// geodesic_acc_core<SPIN, DivFast, FIRST> for two rays
template <bool SPIN, bool FIRST>
__device__ __forceinline__ V3x2 acc2(const Consts& C, const V3x2& q, const V3x2& v, F2 r2, F2 r) {
    const F2 lx = fma2(q.y, v.z, neg2(mul2(q.z, v.y)));
    const F2 ly = fma2(q.z, v.x, neg2(mul2(q.x, v.z)));
    const F2 lz = fma2(q.x, v.y, neg2(mul2(q.y, v.x)));
    V3x2 L;
    L.x = lx; L.y = ly; L.z = lz;
    const F2 L2 = dot3_2(L, L);
    const F2 m = div2(mul2(bc(C.radial_k), L2), mul2(mul2(r2, r2), r));
    V3x2 a;
    if (SPIN) {
        const F2 s = div2(bc(C.drag_k), mul2(r2, r));
        const F2 nqx = neg2(q.x);
        if (FIRST) {
            a.x = fma2(q.x, m, mul2(q.z, s));
            a.z = fma2(q.z, m, mul2(nqx, s));
        } else {
            a.x = fma2(q.z, s, mul2(q.x, m));
            a.z = fma2(nqx, s, mul2(q.z, m));
        }
        a.y = mul2(q.y, m);
    } else {
        a.x = mul2(q.x, m); a.y = mul2(q.y, m); a.z = mul2(q.z, m);
    }
    return a;
}

// rk4_step_fast for two rays; h / hh / h6 per half (pass broadcasts of uniform values for the constant-step burst).
// rmin: the smallest radius stages 2-4 saw, per half.
template <bool SPIN>
__device__ __forceinline__ void rk4_step2(const Consts& C, V3x2& p, V3x2& v, F2 h, F2 hh, F2 h6, F2 r2_0, F2 r_0, float& rmin_lo,
                                          float& rmin_hi) {
    const V3x2 p0 = p, v0 = v;
    const V3x2 k1 = acc2<SPIN, true>(C, p0, v0, r2_0, r_0);
    const V3x2 v2 = axpy2(v0, k1, hh), p2 = axpy2(p0, v0, hh);
    const F2 r2_2 = dot3_2(p2, p2), r_2 = sqrt2(r2_2);
    const V3x2 k2 = acc2<SPIN, false>(C, p2, v2, r2_2, r_2);
    const V3x2 v3 = axpy2(v0, k2, hh), p3 = axpy2(p0, v2, hh);
    const F2 r2_3 = dot3_2(p3, p3), r_3 = sqrt2(r2_3);
    const V3x2 k3 = acc2<SPIN, false>(C, p3, v3, r2_3, r_3);
    const V3x2 v4 = axpy2(v0, k3, h), p4 = axpy2(p0, v3, h);
    const F2 r2_4 = dot3_2(p4, p4), r_4 = sqrt2(r2_4);
    const V3x2 k4 = acc2<SPIN, false>(C, p4, v4, r2_4, r_4);
    V3x2 sv, sp;
    sv.x = rk4_sum2(k1.x, k2.x, k3.x, k4.x); sv.y = rk4_sum2(k1.y, k2.y, k3.y, k4.y); sv.z = rk4_sum2(k1.z, k2.z, k3.z, k4.z);
    sp.x = rk4_sum2(v0.x, v2.x, v3.x, v4.x); sp.y = rk4_sum2(v0.y, v2.y, v3.y, v4.y); sp.z = rk4_sum2(v0.z, v2.z, v3.z, v4.z);
    v = axpy2(v0, sv, h6);
    p = axpy2(p0, sp, h6);
    float a0, a1, b0, b1, c0, c1;
    upk(r_2, a0, a1); upk(r_3, b0, b1); upk(r_4, c0, c1);
    rmin_lo = fminf(a0, fminf(b0, c0));
    rmin_hi = fminf(a1, fminf(b1, c1));
}
}  // namespace rrtp

namespace {

constexpr int kPTileW = 16, kPTileH = 2 * 32 / kPTileW;   // one warp = 16 x 4 pixels, two per thread
#ifndef RRT_PACKED_MIN_BLOCKS
#define RRT_PACKED_MIN_BLOCKS 16        // one-warp CTAs: <= 128 registers
#endif
#ifndef RRT_PACKED_MIN_BLOCKS_MEDIA
#define RRT_PACKED_MIN_BLOCKS_MEDIA 16
#endif

template <bool SPIN, bool MEDIA>
__global__ void __launch_bounds__(32, MEDIA ? RRT_PACKED_MIN_BLOCKS_MEDIA : RRT_PACKED_MIN_BLOCKS) render_kernel_p(const __grid_constant__ FrameArgs A) {
    using rrtp::F2;
    using rrtp::V3x2;
    const Consts& C = A.C;
    const int lane = threadIdx.x & 31;
    const int ntx = (A.w + kPTileW - 1) / kPTileW;
    const int nty = (A.local_rows + kPTileH - 1) / kPTileH;
    const unsigned ntiles = (unsigned)(ntx * nty);
    const int max_steps = C.max_steps;
    const bool want_disk = (C.flags & RRT_FLAG_DISK) != 0, want_dust = (C.flags & RRT_FLAG_DUST) != 0;
    const float zone_rmax = fmaxf(18.0f, fmaxf(C.disk_zone_r, C.dust_zone_r));
    const V3 cam_p = mk(A.cam.pos[0], A.cam.pos[1], A.cam.pos[2]);
    const bool fast_ok = rrt::dot3(cam_p, cam_p) < 1.0e8f && C.acc_rmin < C.horizon_r && C.horizon_r >= 1e-3f;
    const float redo_below = fast_ok ? C.acc_rmin : __int_as_float(0x7f800000);
    // inert state of a finished half: at rest on the spin axis, L = 0 and s x p = 0, so the RHS is exactly zero and the
    // half sits at r = 100 -- inside the burst window, never captured, never escaping
    const V3 park_p = mk(0.0f, 100.0f, 0.0f), park_v = mk(0.0f, 0.0f, 0.0f);
#if RRT_BURST_K > 0
    constexpr int kBurst = RRT_BURST_K;
    const float burst_min = fmaxf(zone_rmax, fmaxf(C.horizon_r, redo_below));
    const float burst_margin = 1.25f * (float)kBurst * C.h[0];
    const float burst_lo = burst_min + burst_margin, burst_hi = 250.0f - burst_margin;
#endif

    unsigned long long c_steps = 0, c_disk = 0, c_dust = 0, c_dense = 0;
    unsigned c_cap = 0, c_esc = 0, c_exh = 0, c_touch = 0;

    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(A.ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int tx = (int)(tile % (unsigned)ntx), kk = (int)(tile / (unsigned)ntx);
        const int cc = nty >> 1, mm = min(cc, nty - 1 - cc);  // centre-out row order, see render_kernel
        int ty;
        if (kk <= 2 * mm) ty = (kk & 1) ? cc + ((kk + 1) >> 1) : cc - (kk >> 1);
        else ty = (cc > nty - 1 - cc) ? (cc - mm - 1) - (kk - (2 * mm + 1)) : (cc + mm + 1) + (kk - (2 * mm + 1));
        const int x0 = tx * kPTileW + 2 * (lane & 7);
        const int ly = ty * kPTileH + (lane >> 3);
        int y = 0;
        bool row_ok = ly < A.local_rows;
        if (row_ok) {
            const int grp = ly / A.band_group;
            y = (grp * A.band_nranks + A.band_rank) * A.band_group + (ly - grp * A.band_group);
            row_ok = y < A.h;
        }
        bool alive[2] = {row_ok && x0 < A.w, row_ok && x0 + 1 < A.w};
        if (!alive[0] && !alive[1]) continue;
#ifdef RRT_WITH_TILE_LOG
        unsigned long long t_begin = 0;
        if (A.tile_log) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
#endif

        V3x2 P, V;
        {
            const V3 vA = alive[0] ? ray_dir(A, x0, y) : park_v, vB = alive[1] ? ray_dir(A, x0 + 1, y) : park_v;
            const V3 pA = alive[0] ? cam_p : park_p, pB = alive[1] ? cam_p : park_p;
            P.x = rrtp::pk(pA.x, pB.x); P.y = rrtp::pk(pA.y, pB.y); P.z = rrtp::pk(pA.z, pB.z);
            V.x = rrtp::pk(vA.x, vB.x); V.y = rrtp::pk(vA.y, vB.y); V.z = rrtp::pk(vA.z, vB.z);
        }
        float Ir[2] = {0.f, 0.f}, Ig[2] = {0.f, 0.f}, Ib[2] = {0.f, 0.f}, T[2] = {1.0f, 1.0f};
        bool touched[2] = {false, false};
        unsigned n_disk = 0, n_dust = 0, n_dense = 0;
        int most_steps = 0;

        // retire one half: count it, run the epilogue (reference :123-173), park its state
        auto retire = [&](int hf, int steps, unsigned end, float Tend) {
            const V3 p = rrtp::half_of(P, hf), v = rrtp::half_of(V, hf);
            finish_ray(A, x0 + hf, y, ly, Ir[hf], Ig[hf], Ib[hf], Tend, p, v, steps, end | (touched[hf] ? kEndTouched : 0u));
            c_steps += (unsigned)steps;
            most_steps = max(most_steps, steps);
            c_cap += (end & kEndCaptured) ? 1u : 0u;
            c_exh += (end & kEndExhausted) ? 1u : 0u;
            c_esc += (end & (kEndCaptured | kEndExhausted)) ? 0u : 1u;
            c_touch += touched[hf] ? 1u : 0u;
            alive[hf] = false;
            rrtp::set_half(P, hf, park_p);
            rrtp::set_half(V, hf, park_v);
        };

        int it = 0;
        int burst_after = 0;
        F2 R2 = rrtp::norm2_loop_2(P);
        F2 R = rrtp::sqrt2(R2);                                                              // :44
        for (;;) {
#if RRT_BURST_K > 0
            // ---- phase 1: unchecked vacuum bursts, warp-uniform (see render_kernel / trace_ray) -----------------------
            {
                const unsigned in_loop = __activemask();
                const F2 H = rrtp::bc(C.h[0]), HH = rrtp::bc(C.hh[0]), H6 = rrtp::bc(C.h6[0]);
#pragma unroll 1
                for (;;) {
                    float r0, r1;
                    rrtp::upk(R, r0, r1);
                    const bool ok = fminf(r0, r1) >= burst_lo && fmaxf(r0, r1) <= burst_hi && it >= burst_after && it + kBurst < max_steps;
                    if (!(ok && __activemask() == in_loop)) break;
                    const V3x2 Ps = P, Vs = V;
                    const F2 R2s = R2, Rs = R;
                    float mn = fminf(r0, r1), mx = fmaxf(r0, r1);
#pragma unroll 1
                    for (int kb = 0; kb < kBurst; ++kb) {
                        float ma, mb;
                        rrtp::rk4_step2<SPIN>(C, P, V, H, HH, H6, R2, R, ma, mb);            // :64
                        R2 = rrtp::norm2_loop_2(P);
                        R = rrtp::sqrt2(R2);
                        rrtp::upk(R, r0, r1);
                        mn = fminf(mn, fminf(fminf(ma, mb), fminf(r0, r1)));
                        mx = fmaxf(mx, fmaxf(r0, r1));
                    }
                    if (mn >= burst_min && mx <= 250.0f) { it += kBurst; continue; }
                    P = Ps; V = Vs; R2 = R2s; R = Rs;
                    burst_after = it + kBurst;
                    break;
                }
            }
#endif
            // ---- phase 2: checked iterations ----------------------------------------------------------------------
            bool rewind = false;
#pragma unroll 1
            while (it < max_steps) {                                                         // :41
                float r[2];
                rrtp::upk(R, r[0], r[1]);
                bool parked = false;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    if (alive[hf] && r[hf] < C.horizon_r) {                                  // :47-51
                        retire(hf, it, kEndCaptured, 0.0f);
                        parked = true;
                    }
                if (!(alive[0] || alive[1])) break;
                if (parked) {   // the parked half's radius, so that the step below is taken from a consistent state
                    R2 = rrtp::norm2_loop_2(P);
                    R = rrtp::sqrt2(R2);
                    rrtp::upk(R, r[0], r[1]);
                }
                float h[2], h6[2];
                unsigned zones[2];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    h[hf] = C.h[0];
                    h6[hf] = C.h6[0];
                    zones[hf] = 0u;
                    if (r[hf] < zone_rmax) {
                        const float py = rrtp::half_of(P.y, hf);
                        const bool near_bh = r[hf] < 18.0f;                                  // :56
                        const bool disk_zone = fabsf(py) < C.disk_zone_y && r[hf] < C.disk_zone_r;   // :57
                        const bool dust_zone = fabsf(py) < C.dust_zone_y && r[hf] < C.dust_zone_r;   // :58
                        const int zi = near_bh ? 1 : (disk_zone ? 2 : (dust_zone ? 3 : 0));  // :60-62
                        h[hf] = C.h[zi];
                        h6[hf] = C.h6[zi];
                        if (MEDIA) zones[hf] = (disk_zone && want_disk ? 1u : 0u) | (dust_zone && want_dust ? 2u : 0u);   // :67
                    }
                }
                const F2 H = rrtp::pk(h[0], h[1]), H6 = rrtp::pk(h6[0], h6[1]);
                const F2 HH = rrtp::mul2(H, rrtp::bc(0.5f));   // exact (power of two)
                const V3x2 Q = P, Vin = V;         // pre-step state: media and the escape test use Q (:68-69, :120)
                float rmin[2];
                rrtp::rk4_step2<SPIN>(C, P, V, H, HH, H6, R2, R, rmin[0], rmin[1]);          // :64
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    if (alive[hf] && !(rmin[hf] >= redo_below)) {   // general-domain redo, see trace_ray
                        const PV s = rk4_step_general<SPIN>(C, rrtp::half_of(Q, hf), rrtp::half_of(Vin, hf), h[hf], h[hf] * 0.5f, h6[hf]);
                        rrtp::set_half(P, hf, s.p);
                        rrtp::set_half(V, hf, s.v);
                    }
                ++it;
                const F2 R2n = rrtp::norm2_loop_2(P);
                const F2 Rn = rrtp::sqrt2(R2n);
                if (MEDIA) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf)
                        if (alive[hf] && zones[hf]) {                                        // :67
                            n_disk += zones[hf] & 1u;
                            n_dust += zones[hf] >> 1;
                            const MediaOut m = media_sample(C, rrtp::half_of(Q, hf), rrtp::half_of(V, hf), r[hf], h[hf], A.time, zones[hf]);
                            if (m.dense) {                                                   // :71
                                touched[hf] = true;
                                ++n_dense;
                                const float wgt = rrt::mul(rrt::sub(1.0f, m.s), T[hf]);      // :109
                                Ir[hf] = rrt::mad(m.er, wgt, Ir[hf]); Ig[hf] = rrt::mad(m.eg, wgt, Ig[hf]); Ib[hf] = rrt::mad(m.eb, wgt, Ib[hf]);
                                T[hf] = rrt::mul(T[hf], m.s);                                // :115
                            }
                        }
                }
                bool gone = false;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    if (alive[hf] && r[hf] > 250.0f && rrt::dot3(rrtp::half_of(Q, hf), rrtp::half_of(V, hf)) > 0.0f) {   // :120
                        retire(hf, it, 0u, T[hf]);
                        gone = true;
                    }
                if (!(alive[0] || alive[1])) break;
                if (gone) {
                    R2 = rrtp::norm2_loop_2(P);
                    R = rrtp::sqrt2(R2);
                } else {
                    R2 = R2n;
                    R = Rn;
                }
#if RRT_BURST_K > 0
                {
                    const unsigned in_loop = __activemask();
                    float r0, r1;
                    rrtp::upk(R, r0, r1);
                    if (fminf(r0, r1) >= burst_lo && fmaxf(r0, r1) <= burst_hi && it >= burst_after && it + kBurst < max_steps &&
                        __activemask() == in_loop) { rewind = true; break; }
                }
#endif
            }
            if (!rewind) break;
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
            if (alive[hf]) retire(hf, it, kEndExhausted, T[hf]);  // the loop ran out (:41)
        c_disk += n_disk; c_dust += n_dust; c_dense += n_dense;
#ifdef RRT_WITH_TILE_LOG
        if (A.tile_log) {
            const int most = __reduce_max_sync(__activemask(), most_steps);
            if (lane == __ffs(__activemask()) - 1 && tile < A.tile_log_cap) {
                unsigned long long t_end;
                unsigned smid;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                unsigned long long* e = A.tile_log + 4ull * tile;
                e[0] = t_begin; e[1] = t_end; e[2] = ((unsigned long long)ty << 32) | (unsigned)tx; e[3] = ((unsigned long long)smid << 32) | (unsigned)most;
            }
        }
#endif
    }

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_steps += __shfl_xor_sync(0xffffffffu, c_steps, o);
        c_disk += __shfl_xor_sync(0xffffffffu, c_disk, o);
        c_dust += __shfl_xor_sync(0xffffffffu, c_dust, o);
        c_dense += __shfl_xor_sync(0xffffffffu, c_dense, o);
        c_cap += __shfl_xor_sync(0xffffffffu, c_cap, o);
        c_esc += __shfl_xor_sync(0xffffffffu, c_esc, o);
        c_exh += __shfl_xor_sync(0xffffffffu, c_exh, o);
        c_touch += __shfl_xor_sync(0xffffffffu, c_touch, o);
    }
    if (lane == 0 && A.counters) {
        atomicAdd(A.counters + 0, c_steps);
        atomicAdd(A.counters + 1, c_disk);
        atomicAdd(A.counters + 2, c_dust);
        atomicAdd(A.counters + 3, c_dense);
        atomicAdd(A.counters + 4, (unsigned long long)c_cap);
        atomicAdd(A.counters + 5, (unsigned long long)c_esc);
        atomicAdd(A.counters + 6, (unsigned long long)c_exh);
        atomicAdd(A.counters + 7, (unsigned long long)c_touch);
    }
}

}  // namespace
