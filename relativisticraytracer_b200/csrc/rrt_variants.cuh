// rrt_variants.cuh -- two measured alternatives to render_kernel, kept selectable (RRT_KERNEL_VARIANT=2 / 3) as
// evidence for the design choice; both are bit-identical to render_kernel in the strict contract and neither is
// faster on B200 (profiles/r1_history.md).  Included by rrt_b200.cu after rrt_kernel.cuh, whose helpers they share.
//   render_kernel2  two rays per thread in packed f32x2 registers (FFMA2 / FMUL2 / FADD2)
//   render_kernel3  a wavefront inside every persistent warp: lane refill + compacted media job ring
#pragma once
#include "rrt_kernel.cuh"

namespace rrt {
// ====================================================================================================
// Packed FP32 (Blackwell f32x2): two rays per thread.
//
// sm_100 adds add/sub/mul/fma .f32x2 (SASS FADD2 / FMUL2 / FFMA2): one instruction, two independent IEEE
// binary32 operations on a 64-bit register pair.  Measured on B200 (tools/microbench/f32x2.cu): an FFMA2
// occupies the FMA pipe for two cycles but only ONE issue slot, i.e. the same FLOP rate as scalar FFMA at
// half the instruction count.  The scalar step loop is issue-slot bound (1 warp instruction per SMSP per
// clock, ~20 % of them not FMA-pipe work), so carrying two rays per thread as the two halves of f32x2
// registers removes the issue-slot limit.  Each half is still rounded separately, in the same order as the
// scalar code above: results are bit-identical per ray (all GPU parity tests pass with either kernel).
// What it does NOT remove is the register-file operand bandwidth: an FFMA2 with three distinct register
// pairs costs 3.1 cycles, with two pairs + a broadcast scalar 2.3, an FADD2 2.1 (tools/microbench/
// f32x2_operands.cu), so the packed loop lands at ~330 cycles per ray-step -- the same as the scalar loop,
// which is why the scalar kernel stays the default and this one is kept as a measured alternative.
// ====================================================================================================
typedef unsigned long long F2;  // .lo = ray A, .hi = ray B

__device__ __forceinline__ F2 pk(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(F2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ float lo_of(F2 a) { float l, h; upk(a, l, h); return l; }
__device__ __forceinline__ float hi_of(F2 a) { float l, h; upk(a, l, h); return h; }
__device__ __forceinline__ float half_of(F2 a, int hf) { float l, h; upk(a, l, h); return hf ? h : l; }
__device__ __forceinline__ F2 bc(float c) { return pk(c, c); }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ F2 sub2(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ F2 mul2_raw(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false (it does not do that
// to the scalar .rn forms), and it folds fma(a,b,-0) and fma(a,1,b) back into mul / add first.  A rounded
// product that must stay a rounded product is therefore written as fma(a, b, nz) with nz = -0.0f read from
// the kernel parameters: a*b + (-0) rounds exactly like a*b (also for zero and subnormal products), and
// ptxas cannot fold an addend it does not know.  Same instruction count (FFMA2 instead of FMUL2).
__device__ __forceinline__ F2 mul2(F2 a, F2 b, F2 nz) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz)); return r; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

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

// loop-invariant packed constants
struct K2 {
    F2 one, two, mhalf, zero;
    F2 nz;     // -0.0f from the parameter block (see mul2)
    F2 nk;     // -radial_k
    F2 ndrag;  // -drag_k
};
__device__ __forceinline__ K2 make_k2(const Consts& C) {
    K2 k;
    k.one = bc(1.0f); k.two = bc(2.0f); k.mhalf = bc(-0.5f); k.zero = bc(0.0f);
    k.nz = bc(C.neg_zero);
    k.nk = bc(-C.radial_k);
    k.ndrag = bc(-C.drag_k);
    return k;
}

// x / y for both halves given xneg = -x: the div_rn_fast sequence with the signs moved so that no operand
// negation is needed (IEEE round-to-nearest is sign-symmetric, so every intermediate is the exact negation
// of, or equal to, its scalar counterpart):  nr = rcp(-y) = -r;  e = 1 + y*nr;  nr1 = nr + nr*e = -r1;
// q0 = xneg*nr1 = x*r1;  remn = y*q0 + xneg = -(x - y*q0);  q = q0 + nr1*remn = q0 + r1*rem.
__device__ __forceinline__ F2 div2_negnum(F2 xneg, F2 y, const K2& k) {
    float yl, yh;
    upk(y, yl, yh);
    const F2 nr = pk(rcp_approx(-yl), rcp_approx(-yh));
    const F2 e = fma2(y, nr, k.one);
    const F2 nr1 = fma2(nr, e, nr);
    const F2 q0 = mul2(xneg, nr1, k.nz);
    const F2 remn = fma2(y, q0, xneg);
    return fma2(nr1, remn, q0);
}
// sqrt(x) for both halves: sqrt_rn_fast with g*g - x and -0.5*y instead of x - g*g and 0.5*y
__device__ __forceinline__ F2 sqrt2(F2 x, const K2& k) {
    float xl, xh;
    upk(x, xl, xh);
    const F2 y = pk(rsqrt_approx(xl), rsqrt_approx(xh));
    const F2 g = mul2(x, y, k.nz);
    const F2 nh = mul2(y, k.mhalf, k.nz);
    const F2 en = fma2(g, g, sub2(k.zero, x));
    return fma2(en, nh, g);
}
__device__ __forceinline__ F2 dot2(const V3x2& a, const V3x2& b, const K2& k) { return add2(add2(mul2(a.x, b.x, k.nz), mul2(a.y, b.y, k.nz)), mul2(a.z, b.z, k.nz)); }

// geodesic_acc_fast for two rays
template <bool SPIN>
__device__ __forceinline__ V3x2 geodesic_acc2(const K2& k, const V3x2& q, const V3x2& v, F2 r2, F2 r) {
    const F2 lx = sub2(mul2(q.y, v.z, k.nz), mul2(q.z, v.y, k.nz));
    const F2 ly = sub2(mul2(q.z, v.x, k.nz), mul2(q.x, v.z, k.nz));
    const F2 lz = sub2(mul2(q.x, v.y, k.nz), mul2(q.y, v.x, k.nz));
    const F2 L2 = add2(add2(mul2(lx, lx, k.nz), mul2(ly, ly, k.nz)), mul2(lz, lz, k.nz));
    const F2 m = div2_negnum(mul2(k.nk, L2, k.nz), mul2(mul2(r2, r2, k.nz), r, k.nz), k);
    V3x2 a;
    a.x = mul2(q.x, m, k.nz);
    a.y = mul2(q.y, m, k.nz);
    a.z = mul2(q.z, m, k.nz);
    if (SPIN) {
        const F2 s = div2_negnum(k.ndrag, mul2(r2, r, k.nz), k);
        a.x = add2(a.x, mul2(q.z, s, k.nz));
        a.z = sub2(a.z, mul2(q.x, s, k.nz));
    }
    return a;
}

// rk4_step_fast for two rays.  h/hh/h6 are per-half.  rmin_{lo,hi}: smallest radius seen by stages 2-4.
template <bool SPIN>
__device__ __forceinline__ void rk4_step2(const K2& k, V3x2& p, V3x2& v, F2 h, F2 hh, F2 h6, F2 r2_0, F2 r_0, float& rmin_lo,
                                          float& rmin_hi) {
    const V3x2 p0 = p, v0 = v;
    const V3x2 k1 = geodesic_acc2<SPIN>(k, p0, v0, r2_0, r_0);
    V3x2 v2, p2;
    v2.x = add2(v0.x, mul2(k1.x, hh, k.nz)); v2.y = add2(v0.y, mul2(k1.y, hh, k.nz)); v2.z = add2(v0.z, mul2(k1.z, hh, k.nz));
    p2.x = add2(p0.x, mul2(v0.x, hh, k.nz)); p2.y = add2(p0.y, mul2(v0.y, hh, k.nz)); p2.z = add2(p0.z, mul2(v0.z, hh, k.nz));
    const F2 r2_2 = dot2(p2, p2, k), r_2 = sqrt2(r2_2, k);
    const V3x2 k2 = geodesic_acc2<SPIN>(k, p2, v2, r2_2, r_2);
    V3x2 v3, p3;
    v3.x = add2(v0.x, mul2(k2.x, hh, k.nz)); v3.y = add2(v0.y, mul2(k2.y, hh, k.nz)); v3.z = add2(v0.z, mul2(k2.z, hh, k.nz));
    p3.x = add2(p0.x, mul2(v2.x, hh, k.nz)); p3.y = add2(p0.y, mul2(v2.y, hh, k.nz)); p3.z = add2(p0.z, mul2(v2.z, hh, k.nz));
    const F2 r2_3 = dot2(p3, p3, k), r_3 = sqrt2(r2_3, k);
    const V3x2 k3 = geodesic_acc2<SPIN>(k, p3, v3, r2_3, r_3);
    V3x2 v4, p4;
    v4.x = add2(v0.x, mul2(k3.x, h, k.nz)); v4.y = add2(v0.y, mul2(k3.y, h, k.nz)); v4.z = add2(v0.z, mul2(k3.z, h, k.nz));
    p4.x = add2(p0.x, mul2(v3.x, h, k.nz)); p4.y = add2(p0.y, mul2(v3.y, h, k.nz)); p4.z = add2(p0.z, mul2(v3.z, h, k.nz));
    const F2 r2_4 = dot2(p4, p4, k), r_4 = sqrt2(r2_4, k);
    const V3x2 k4 = geodesic_acc2<SPIN>(k, p4, v4, r2_4, r_4);
    // k1 + (2*k2 + (2*k3 + k4)); 2*x exact => fused form rounds identically (see rk4_step)
    const F2 svx = add2(k1.x, fma2(k.two, k2.x, fma2(k.two, k3.x, k4.x)));
    const F2 svy = add2(k1.y, fma2(k.two, k2.y, fma2(k.two, k3.y, k4.y)));
    const F2 svz = add2(k1.z, fma2(k.two, k2.z, fma2(k.two, k3.z, k4.z)));
    const F2 spx = add2(v0.x, fma2(k.two, v2.x, fma2(k.two, v3.x, v4.x)));
    const F2 spy = add2(v0.y, fma2(k.two, v2.y, fma2(k.two, v3.y, v4.y)));
    const F2 spz = add2(v0.z, fma2(k.two, v2.z, fma2(k.two, v3.z, v4.z)));
    v.x = add2(v0.x, mul2(svx, h6, k.nz)); v.y = add2(v0.y, mul2(svy, h6, k.nz)); v.z = add2(v0.z, mul2(svz, h6, k.nz));
    p.x = add2(p0.x, mul2(spx, h6, k.nz)); p.y = add2(p0.y, mul2(spy, h6, k.nz)); p.z = add2(p0.z, mul2(spz, h6, k.nz));
    float a0, a1, b0, b1, c0, c1;
    upk(r_2, a0, a1); upk(r_3, b0, b1); upk(r_4, c0, c1);
    rmin_lo = fminf(a0, fminf(b0, c0));
    rmin_hi = fminf(a1, fminf(b1, c1));
}

}  // namespace rrt

namespace {

// =====================================================================================================
// render_kernel2 (opt-in, RRT_KERNEL_VARIANT=2; bit-identical output, same speed as render_kernel on B200 --
// see profiles/r1_history.md "packed f32x2 experiment"):
// two rays per thread in packed f32x2 registers (see include/rrt_device.cuh, "Packed FP32").
// One warp = a 16x4-pixel tile; thread (lx, ly) owns pixels (2*lx, ly) and (2*lx+1, ly) of the tile.  The
// two rays step in lock-step (same iteration index); a ray that terminates is finalised at once (sky,
// effects, store) and its half is parked on a harmless far-away state until its partner finishes.
// =====================================================================================================
constexpr int kTile2W = 16;

template <bool SPIN, bool MEDIA>
__global__ void __launch_bounds__(kBlock, RRT_MIN_BLOCKS) render_kernel2(const __grid_constant__ FrameArgs A) {
    using rrt::F2;
    using rrt::V3x2;
    const Consts& C = A.C;
    const int lane = threadIdx.x & 31;
    const int ntx = (A.w + kTile2W - 1) / kTile2W;
    const int nty = (A.local_rows + kTileH - 1) / kTileH;
    const unsigned ntiles = (unsigned)(ntx * nty);
    const rrt::K2 k2 = rrt::make_k2(C);
    const F2 kHalf = rrt::bc(0.5f);
    const int max_steps = C.max_steps;
    const bool want_disk = (C.flags & RRT_FLAG_DISK) != 0, want_dust = (C.flags & RRT_FLAG_DUST) != 0;
    const float zone_rmax = fmaxf(18.0f, fmaxf(C.disk_zone_r, C.dust_zone_r));
    const V3 cam_p = mk(A.cam.pos[0], A.cam.pos[1], A.cam.pos[2]);
    const bool fast_ok = rrt::dot3(cam_p, cam_p) < 1.0e8f && C.acc_rmin < C.horizon_r && C.horizon_r >= 1e-3f;
    const V3 park_p = mk(1000.0f, 0.0f, 0.0f), park_v = mk(0.0f, 0.0f, 0.0f);  // inert state of a finished half

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
        const int x0 = tx * kTile2W + 2 * (lane & 7);
        const int ly = ty * kTileH + (lane >> 3);
        int y = 0;
        bool row_ok = ly < A.local_rows;
        if (row_ok) {
            const int grp = ly / A.band_group;
            y = (grp * A.band_nranks + A.band_rank) * A.band_group + (ly - grp * A.band_group);
            row_ok = y < A.h;
        }
        bool alive[2] = {row_ok && x0 < A.w, row_ok && x0 + 1 < A.w};
        if (!alive[0] && !alive[1]) continue;

        V3x2 P, V;
        {
            const V3 vA = alive[0] ? ray_dir(A, x0, y) : park_v, vB = alive[1] ? ray_dir(A, x0 + 1, y) : park_v;
            const V3 pA = alive[0] ? cam_p : park_p, pB = alive[1] ? cam_p : park_p;
            P.x = rrt::pk(pA.x, pB.x); P.y = rrt::pk(pA.y, pB.y); P.z = rrt::pk(pA.z, pB.z);
            V.x = rrt::pk(vA.x, vB.x); V.y = rrt::pk(vA.y, vB.y); V.z = rrt::pk(vA.z, vB.z);
        }
        float Ir[2] = {0.f, 0.f}, Ig[2] = {0.f, 0.f}, Ib[2] = {0.f, 0.f}, T[2] = {1.0f, 1.0f};
        bool touched[2] = {false, false};
        unsigned n_disk = 0, n_dust = 0, n_dense = 0;

        // retire one half: count it, run the epilogue, park its state
        auto retire = [&](int hf, int steps, unsigned end, float Tend) {
            const V3 p = rrt::half_of(P, hf), v = rrt::half_of(V, hf);
            finish_ray(A, x0 + hf, y, ly, Ir[hf], Ig[hf], Ib[hf], Tend, p, v, steps, end | (touched[hf] ? kEndTouched : 0u));
            c_steps += (unsigned)steps;
            c_cap += (end & kEndCaptured) ? 1u : 0u;
            c_exh += (end & kEndExhausted) ? 1u : 0u;
            c_esc += (end & (kEndCaptured | kEndExhausted)) ? 0u : 1u;
            c_touch += touched[hf] ? 1u : 0u;
            alive[hf] = false;
            rrt::set_half(P, hf, park_p);
            rrt::set_half(V, hf, park_v);
        };

        int it = 0;
#pragma unroll 1
        for (; it < max_steps; ++it) {                                                       // :41
            F2 R2 = rrt::dot2(P, P, k2);
            F2 R = rrt::sqrt2(R2, k2);                                                       // :44
            float r[2];
            rrt::upk(R, r[0], r[1]);
            bool parked = false;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
                if (alive[hf] && r[hf] < C.horizon_r) {                                      // :47-51
                    retire(hf, it, kEndCaptured, 0.0f);
                    r[hf] = 1000.0f;
                    parked = true;
                }
            if (!(alive[0] || alive[1])) break;
            if (parked) {  // radius of the parked half only, so the step below stays in the fast domain
                float r2l, r2h;
                rrt::upk(R2, r2l, r2h);
                R = rrt::pk(r[0], r[1]);
                R2 = rrt::pk(alive[0] ? r2l : 1.0e6f, alive[1] ? r2h : 1.0e6f);
            }
            float h[2], h6[2];
            unsigned zones[2];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                h[hf] = C.h[0];
                h6[hf] = C.h6[0];
                zones[hf] = 0u;
                if (r[hf] < zone_rmax) {
                    const float py = rrt::half_of(P.y, hf);
                    const bool near_bh = r[hf] < 18.0f;                                      // :56
                    const bool disk_zone = fabsf(py) < C.disk_zone_y && r[hf] < C.disk_zone_r;   // :57
                    const bool dust_zone = fabsf(py) < C.dust_zone_y && r[hf] < C.dust_zone_r;   // :58
                    const int zi = near_bh ? 1 : (disk_zone ? 2 : (dust_zone ? 3 : 0));      // :60-62
                    h[hf] = C.h[zi];
                    h6[hf] = C.h6[zi];
                    zones[hf] = (disk_zone && want_disk ? 1u : 0u) | (dust_zone && want_dust ? 2u : 0u);
                }
            }
            const F2 H = rrt::pk(h[0], h[1]), H6 = rrt::pk(h6[0], h6[1]);
            const F2 HH = rrt::mul2_raw(H, kHalf);  // exact (power of two)
            const V3x2 Q = P, Vin = V;          // pre-step state: media and the escape test use Q (:68-69, :120)
            float rmin[2];
            rrt::rk4_step2<SPIN>(k2, P, V, H, HH, H6, R2, R, rmin[0], rmin[1]);              // :64
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
                if (alive[hf] && (!fast_ok || rmin[hf] < C.acc_rmin)) {  // general-domain redo, see render_kernel
                    const PV s = rk4_step_general<SPIN>(C, rrt::half_of(Q, hf), rrt::half_of(Vin, hf), h[hf], h[hf] * 0.5f, h6[hf]);
                    rrt::set_half(P, hf, s.p);
                    rrt::set_half(V, hf, s.v);
                }
            if (MEDIA) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    if (alive[hf] && zones[hf]) {                                            // :67
                        n_disk += zones[hf] & 1u;
                        n_dust += zones[hf] >> 1;
                        const MediaOut m = media_sample(C, rrt::half_of(Q, hf), rrt::half_of(V, hf), r[hf], h[hf], A.time, zones[hf]);
                        if (m.dense) {                                                       // :71
                            touched[hf] = true;
                            ++n_dense;
                            const float wgt = (1.0f - m.s) * T[hf];                          // :109
                            Ir[hf] += m.er * wgt; Ig[hf] += m.eg * wgt; Ib[hf] += m.eb * wgt;    // :111-113
                            T[hf] *= m.s;                                                    // :115
                        }
                    }
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
                if (alive[hf] && r[hf] > 250.0f && rrt::dot3(rrt::half_of(Q, hf), rrt::half_of(V, hf)) > 0.0f)   // :120
                    retire(hf, it + 1, 0u, T[hf]);
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
            if (alive[hf]) retire(hf, it, kEndExhausted, T[hf]);  // the loop ran out (:41)
        c_disk += n_disk; c_dust += n_dust; c_dense += n_dense;
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

// =====================================================================================================
// render_kernel3 (opt-in, RRT_KERNEL_VARIANT=3; bit-identical output, slower than render_kernel at N=1 -- see
// profiles/r1_history.md "wavefront-in-a-warp experiment"): a wavefront inside every persistent warp.
//
// What limits render_kernel on B200 is not the vacuum step (straight-line FMA code, all lanes busy) but the
// side work: media samples evaluated under per-lane branches, and the fact that one 8x4 tile of disk-plane
// rays is ~5e6 warp-instructions of sequential work, which is what a band-parallel 8-GPU frame ends up
// waiting for.  Here:
//   * every lane owns one ray and is refilled individually from a global ticket when its ray ends
//     (persistent threads with lane refill).  Tickets are 4-pixel strips and 8 consecutive strips come from
//     8 distant parts of the centre-out ordered frame, so the expensive disk-plane rays are spread four to
//     a warp instead of 32 to a warp;
//   * a lane that is inside a medium does not evaluate it: it appends a SAMPLE JOB (pre-step position,
//     post-step velocity, zone bits, step-size index, owner lane) to a per-warp ring in shared memory and
//     keeps stepping.  Every kBurst iterations the warp pumps the ring through three stages, each run with
//     one job per lane as soon as 32 jobs are waiting for it:
//        1. disk density + dust envelope of 32 consecutive jobs (survivors of the envelope test are listed),
//        2. the domain-warped ridge noise of 32 listed dust jobs,
//        3. redshift / emission / exp(-tau) of 32 consecutive completed jobs, after which every owner lane
//           folds the results of ITS jobs into its (I, T) registers in ring order.
//     The ring is FIFO and a ray's jobs are appended in step order, so each ray sees exactly the
//     reference's sequence of I += e(1-s)T; T *= s updates (raymarcher.cu:107-115): results are
//     bit-identical to render_kernel;
//   * a ray that ends with jobs still in the ring forces a drain first; the capture rule T = 0
//     (raymarcher.cu:49) is applied when the ray is finalised, after its queued emission has been added with
//     the running T.
// =====================================================================================================
constexpr int kStripW = 4;       // pixels per ticket strip
constexpr int kInterleave = 8;   // consecutive strips are taken from this many distant parts of the frame
constexpr int kRing = 256;       // sample-job slots per warp (power of two)
constexpr int kBurst = 2;        // loop iterations between two control points
constexpr int kRetireMin = 4;    // finished lanes wait until this many can be finalised + refilled together

struct WarpQueue {
    float qx[kRing], qy[kRing], qz[kRing], vx[kRing], vy[kRing], vz[kRing];
    float dd[kRing], dc[kRing];   // densities, filled by stages 1 and 2 (dc holds the dust envelope in between)
    unsigned meta[kRing];         // owner lane | zones << 5 | step-size index << 7
    unsigned dlist[kRing];        // ring of job sequence numbers waiting for stage 2
    float er[32], eg[32], eb[32], s[32];  // stage-3 results of the batch being folded
    unsigned tail;                // sequence number of the next job
};

// ticket -> pixel.  Tickets count pixels of 4x1 strips; strip s is strip (s % 8) * part + s / 8 of the base order,
// the base order being rows from the band centre outwards (see render_kernel), left to right.
__device__ __forceinline__ bool ticket_pixel(const FrameArgs& A, unsigned idx, int spr, unsigned nstrips, unsigned part, int& x,
                                             int& ly) {
    const unsigned s = idx / kStripW, l = idx - s * kStripW;
    const unsigned j = s % kInterleave, b = j * part + s / kInterleave;
    if (b >= nstrips) return false;
    const int k = (int)(b / (unsigned)spr), sx = (int)(b - (unsigned)k * (unsigned)spr);
    const int rows = A.local_rows, c = rows >> 1, m = min(c, rows - 1 - c);
    if (k <= 2 * m) ly = (k & 1) ? c + ((k + 1) >> 1) : c - (k >> 1);
    else ly = (c > rows - 1 - c) ? (c - m - 1) - (k - (2 * m + 1)) : (c + m + 1) + (k - (2 * m + 1));
    x = sx * kStripW + (int)l;
    return x < A.w;
}
__device__ __forceinline__ int band_row(const FrameArgs& A, int ly) {
    const int grp = ly / A.band_group;
    return (grp * A.band_nranks + A.band_rank) * A.band_group + (ly - grp * A.band_group);
}

template <bool SPIN, bool MEDIA>
__global__ void __launch_bounds__(kBlock, RRT_MIN_BLOCKS) render_kernel3(const __grid_constant__ FrameArgs A) {
    __shared__ WarpQueue s_queue[MEDIA ? kBlock / 32 : 1];
    WarpQueue& Q = s_queue[MEDIA ? (threadIdx.x >> 5) : 0];
    const Consts& C = A.C;
    const unsigned FULL = 0xffffffffu, RM = kRing - 1;
    const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const int spr = (A.w + kStripW - 1) / kStripW;
    const unsigned nstrips = (unsigned)spr * (unsigned)A.local_rows;
    const unsigned part = (nstrips + kInterleave - 1) / kInterleave;
    const unsigned total = part * kInterleave * kStripW;
    const int max_steps = C.max_steps;
    const bool want_disk = (C.flags & RRT_FLAG_DISK) != 0, want_dust = (C.flags & RRT_FLAG_DUST) != 0;
    const float zone_rmax = fmaxf(18.0f, fmaxf(C.disk_zone_r, C.dust_zone_r));
    const V3 cam_p = mk(A.cam.pos[0], A.cam.pos[1], A.cam.pos[2]);
    const bool fast_ok = rrt::dot3(cam_p, cam_p) < 1.0e8f && C.acc_rmin < C.horizon_r && C.horizon_r >= 1e-3f;

    unsigned long long c_steps = 0;
    unsigned c_disk = 0, c_dust = 0, c_dense = 0, c_cap = 0, c_esc = 0, c_exh = 0, c_touch = 0;

    enum : int { kEmpty = 0, kActive = 1, kDone = 2 };
    int st = kEmpty;
    unsigned end = 0;      // kEnd* bits of the current ray
    unsigned npend = 0;    // this lane's jobs not yet folded
    V3 p = cam_p, v = mk(0.f, 0.f, 0.f);
    float Ir = 0.f, Ig = 0.f, Ib = 0.f, T = 1.0f;
    int it = 0, x = 0, ly = 0;
    // warp-uniform ring state (sequence numbers; slot = seq & RM):  head <= s1 <= tail
    unsigned head = 0;     // oldest job not yet folded
    unsigned s1 = 0;       // next job for stage 1
    unsigned dhead = 0, dtail = 0;  // stage-2 list
    unsigned wit = 0;      // control points seen
    bool tickets_left = true;
    if (MEDIA) {
        if (lane == 0) Q.tail = 0u;
        __syncwarp();
    }

    // Run every stage that has a full batch; with drain = true run them until the ring is empty.
    auto pump = [&](bool drain) {
        __syncwarp();
        const unsigned tail = *(volatile unsigned*)&Q.tail;
        // ---- stage 1: disk density (:68) and dust envelope (densities.h:70-84) ----
        while (tail - s1 >= 32u || (drain && tail != s1)) {
            const unsigned n = min(32u, tail - s1);
            bool need = false;
            unsigned seq = s1 + lane;
            if (lane < n) {
                const unsigned sl = seq & RM, m = Q.meta[sl];
                const V3 jq = mk(Q.qx[sl], Q.qy[sl], Q.qz[sl]);
                Q.dd[sl] = (m & 32u) ? rrt::disk_density(C, jq, A.time) : 0.0f;
                const float base = (m & 64u) ? rrt::dust_base(C, jq) : 0.0f;
                Q.dc[sl] = base;
                need = base != 0.0f;
            }
            const unsigned nm = __ballot_sync(FULL, need);
            if (need) Q.dlist[(dtail + (unsigned)__popc(nm & lt_mask)) & RM] = seq;
            dtail += (unsigned)__popc(nm);
            s1 += n;
            __syncwarp();
        }
        // ---- stage 2: dust strands (densities.h:86-131) ----
        while (dtail - dhead >= 32u || (drain && dtail != dhead)) {
            const unsigned n = min(32u, dtail - dhead);
            if (lane < n) {
                const unsigned sl = Q.dlist[(dhead + lane) & RM] & RM;
                const V3 jq = mk(Q.qx[sl], Q.qy[sl], Q.qz[sl]);
                Q.dc[sl] = rrt::dust_strands(C, jq, A.time, Q.dc[sl]);                        // :69
            }
            dhead += n;
            __syncwarp();
        }
        // ---- stage 3: transfer of completed jobs, in ring order ----
        const unsigned ready = (dhead == dtail) ? s1 : Q.dlist[dhead & RM];  // first job still waiting for stage 2
        while (ready - head >= 32u || (drain && ready != head)) {
            const unsigned n = min(32u, ready - head);
            const bool have = lane < n;
            unsigned owner = 0;
            bool dense = false;
            if (have) {
                const unsigned sl = (head + lane) & RM, m = Q.meta[sl];
                owner = m & 31u;
                const V3 jq = mk(Q.qx[sl], Q.qy[sl], Q.qz[sl]);
                const V3 jv = mk(Q.vx[sl], Q.vy[sl], Q.vz[sl]);
                const float jr = rrt::sqrt_rn_fast(rrt::dot3(jq, jq));  // the loop header's r of that step (:43-44)
                const unsigned zi = m >> 7;
                const float jh = zi == 1u ? C.h[1] : (zi == 2u ? C.h[2] : (zi == 3u ? C.h[3] : C.h[0]));
                const MediaOut o = media_final(C, jq, jv, jr, jh, Q.dd[sl], Q.dc[sl]);
                dense = o.dense != 0;
                Q.er[lane] = o.er; Q.eg[lane] = o.eg; Q.eb[lane] = o.eb; Q.s[lane] = o.s;
            }
            const unsigned dense_m = __ballot_sync(FULL, dense);
            unsigned mine = n >= 32u ? FULL : ((1u << n) - 1u);  // becomes: batch entries owned by this lane
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const unsigned bk = __ballot_sync(FULL, have && ((owner >> k) & 1u));
                mine &= ((lane >> k) & 1u) ? bk : ~bk;
            }
            __syncwarp();
            npend -= (unsigned)__popc(mine);
            mine &= dense_m;                                                                  // :71
            while (mine) {
                const int j = __ffs((int)mine) - 1;
                mine &= mine - 1u;
                const float s = Q.s[j];
                const float wgt = (1.0f - s) * T;                                             // :109
                Ir += Q.er[j] * wgt; Ig += Q.eg[j] * wgt; Ib += Q.eb[j] * wgt;                // :111-113
                T *= s;                                                                       // :115
                end |= kEndTouched;
                ++c_dense;
            }
            head += n;
            __syncwarp();
        }
    };

    for (;;) {
        // ---- control point: finalise finished rays, refill empty lanes ----
        const unsigned act_m = __ballot_sync(FULL, st == kActive);
        const unsigned done_m = __ballot_sync(FULL, st == kDone);
        if (act_m == 0u || (done_m != 0u && (__popc(done_m) >= kRetireMin || (wit & 15u) == 0u))) {
            if (done_m) {
                if (MEDIA && __any_sync(FULL, st == kDone && npend != 0u)) pump(true);
                if (st == kDone) {                                                            // reference :123-173
                    const bool captured = (end & kEndCaptured) != 0;
                    finish_ray(A, x, band_row(A, ly), ly, Ir, Ig, Ib, captured ? 0.0f : T, p, v, it, end);   // T = 0: :49
                    c_steps += (unsigned)it;
                    c_cap += captured ? 1u : 0u;
                    c_exh += (end & kEndExhausted) ? 1u : 0u;
                    c_esc += (end & (kEndCaptured | kEndExhausted)) ? 0u : 1u;
                    c_touch += (end & kEndTouched) ? 1u : 0u;
                    st = kEmpty;
                }
            }
            if (tickets_left) {
                const unsigned empty_m = __ballot_sync(FULL, st == kEmpty);
                const unsigned n = (unsigned)__popc(empty_m);
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(A.ticket, n);
                base = __shfl_sync(FULL, base, 0);
                if (base + n >= total) tickets_left = false;
                if (st == kEmpty) {
                    const unsigned idx = base + (unsigned)__popc(empty_m & lt_mask);
                    if (idx < total && ticket_pixel(A, idx, spr, nstrips, part, x, ly)) {
                        p = cam_p;
                        v = ray_dir(A, x, band_row(A, ly));                                   // :20-34
                        Ir = 0.f; Ig = 0.f; Ib = 0.f; T = 1.0f;                               // :36-38
                        it = 0;
                        end = max_steps > 0 ? 0u : kEndExhausted;
                        st = max_steps > 0 ? kActive : kDone;
                    }
                }
            }
            if (!__any_sync(FULL, st != kEmpty)) {
                if (!tickets_left) break;
                continue;
            }
        }
        ++wit;

        // ---- kBurst loop iterations of reference :41-121 for every active lane, no warp-wide operation inside ----
#pragma unroll 1
        for (int b = 0; b < kBurst; ++b) {
            if (st == kActive) {
                const float r2 = rrt::dot3(p, p);
                const float r = rrt::sqrt_rn_fast(r2);                                        // :44
                if (r < C.horizon_r) {                                                        // :47-51
                    st = kDone;
                    end |= kEndCaptured;
                } else {
                    bool disk_zone = false, dust_zone = false;
                    int zi = 0;
                    float h = C.h[0], h6 = C.h6[0];
                    if (r < zone_rmax) {
                        const bool near_bh = r < 18.0f;                                       // :56
                        disk_zone = fabsf(p.y) < C.disk_zone_y && r < C.disk_zone_r;          // :57
                        dust_zone = fabsf(p.y) < C.dust_zone_y && r < C.dust_zone_r;          // :58
                        zi = near_bh ? 1 : (disk_zone ? 2 : (dust_zone ? 3 : 0));             // :60-62
                        // selects between uniform constants, not an indexed constant load: lanes of one warp are in
                        // different zones here and a divergent c[][] index is replayed per distinct address
                        h = near_bh ? C.h[1] : (disk_zone ? C.h[2] : (dust_zone ? C.h[3] : C.h[0]));
                        h6 = near_bh ? C.h6[1] : (disk_zone ? C.h6[2] : (dust_zone ? C.h6[3] : C.h6[0]));
                    }
                    const float hh = h * 0.5f;  // exact
                    const V3 q = p, v_in = v;
                    const float rmin = rrt::rk4_step_fast<SPIN>(C, p, v, h, hh, h6, r2, r);   // :64
                    if (!fast_ok || rmin < C.acc_rmin) {  // general-domain redo, see render_kernel
                        const PV sres = rk4_step_general<SPIN>(C, q, v_in, h, hh, h6);
                        p = sres.p; v = sres.v;
                    }
                    ++it;
                    if (MEDIA && (disk_zone || dust_zone)) {                                  // :67
                        const unsigned zones = (disk_zone && want_disk ? 1u : 0u) | (dust_zone && want_dust ? 2u : 0u);
                        c_disk += zones & 1u;
                        c_dust += zones >> 1;
                        if (zones) {
                            // both density functions return 0 outside ISCO <= R <= DISK_OUT (densities.h:21-23,
                            // 70-72): only samples inside that ring become jobs
                            const float R = rrt::sqrt_rn_fast(q.x * q.x + 0.0f * 0.0f + q.z * q.z);
                            if (R >= C.isco && R <= C.disk_out) {
                                // warp-aggregated append: one shared-memory atomic per converged group of lanes
                                const unsigned am = __activemask();
                                const int leader = __ffs((int)am) - 1;
                                unsigned seq = 0;
                                if ((int)lane == leader) seq = atomicAdd(&Q.tail, (unsigned)__popc(am));
                                seq = __shfl_sync(am, seq, leader) + (unsigned)__popc(am & lt_mask);
                                const unsigned sl = seq & RM;
                                Q.qx[sl] = q.x; Q.qy[sl] = q.y; Q.qz[sl] = q.z;
                                Q.vx[sl] = v.x; Q.vy[sl] = v.y; Q.vz[sl] = v.z;               // post-step velocity (:77)
                                Q.meta[sl] = lane | (zones << 5) | ((unsigned)zi << 7);
                                ++npend;
                            }
                        }
                    }
                    if (r > 250.0f && rrt::dot3(q, v) > 0.0f) st = kDone;                     // :120
                    else if (it >= max_steps) { st = kDone; end |= kEndExhausted; }           // :41
                }
            }
        }
        if (MEDIA) {
            __syncwarp();
            const unsigned tail = *(volatile unsigned*)&Q.tail;
            if (tail - s1 >= 32u) pump(false);
            if (tail - head > (unsigned)(kRing - 32 * kBurst)) pump(true);  // no room for another burst: drain
        }
    }

    // one set of atomics per warp
    unsigned long long w_disk = c_disk, w_dust = c_dust, w_dense = c_dense;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_steps += __shfl_xor_sync(FULL, c_steps, o);
        w_disk += __shfl_xor_sync(FULL, w_disk, o);
        w_dust += __shfl_xor_sync(FULL, w_dust, o);
        w_dense += __shfl_xor_sync(FULL, w_dense, o);
        c_cap += __shfl_xor_sync(FULL, c_cap, o);
        c_esc += __shfl_xor_sync(FULL, c_esc, o);
        c_exh += __shfl_xor_sync(FULL, c_exh, o);
        c_touch += __shfl_xor_sync(FULL, c_touch, o);
    }
    if (lane == 0 && A.counters) {
        atomicAdd(A.counters + 0, c_steps);
        atomicAdd(A.counters + 1, w_disk);
        atomicAdd(A.counters + 2, w_dust);
        atomicAdd(A.counters + 3, w_dense);
        atomicAdd(A.counters + 4, (unsigned long long)c_cap);
        atomicAdd(A.counters + 5, (unsigned long long)c_esc);
        atomicAdd(A.counters + 6, (unsigned long long)c_exh);
        atomicAdd(A.counters + 7, (unsigned long long)c_touch);
    }
}

}  // namespace
