// rrt_split.cuh -- the split render pipeline: trace / media / fold as three kernels over a sample pool in HBM.
//
// Why.  A media sample (reference src/raymarcher.cu:67-115) does not feed back into the trajectory: the densities are
// taken at the pre-step position, the redshift at the post-step velocity, and the result only enters
// `I += e (1 - s) T; T *= s`.  In the fused render_kernel a ray that stays in the disk plane is nevertheless ONE
// dependency chain of 2000 x (RK4 step + disk density + dust density) ~ 2*10^7 warp instructions: 17 ms on an empty
// GPU, 60-70 ms when its SM sub-partition is shared six ways -- the tail of every frame and the whole single-frame
// latency of a band-parallel frame on 8 GPUs (profiles/r2_tile_timeline_*.txt).  Its media code also runs with the
// lanes the tile happens to have inside a zone (25 of 32 on the headline frame, fewer on mixed tiles), under the
// register cap and instruction-cache pressure of sharing a kernel with the step loop.
//
// What.  Per frame, on one stream, in passes that each fit the pool:
//   trace_kernel   the same persistent one-warp tile loop and the same trace_ray as render_kernel, in kMediaEmit mode:
//                  an in-zone step appends its sample (32 B: q, v, r, zones) to the warp's record stream in the pool
//                  instead of evaluating it.  A tile that emitted nothing is finished on the spot; otherwise the 32
//                  exit states go to the pool too and the tile is queued for the fold.  The longest chain is now the
//                  trajectory alone (2000 x ~300 instructions, < 1 ms).
//   media_kernel   one thread per pool slot, every lane busy: evaluates the sample with the SAME out-of-line functions
//                  the fused kernel calls (disk_density / dust_base / dust_strands / media_final), the expensive dust
//                  strands compacted once more through shared memory, and overwrites the slot with (e.rgb, s).
//   fold_kernel    one warp per queued tile: walks the tile's records in step order, folds `I += e (1 - s) T; T *= s`
//                  with the fused kernel's own operations, then background, effects, tonemap and store.
// Every ray is traced, sampled and folded by the same arithmetic in the same order as in render_kernel: frames,
// planes and counters are bit-identical (tests/test_gpu_split.py).
//
// Pool (one per stream in flight, HBM; `slot` = 32 bytes):
//   record   = 1 header slot {kRec, lane mask} + popc(mask) sample slots, written by the lanes of a warp that emit
//              together; records of a warp follow each other inside the warp's current chunk, chunks are chained by
//              {kJump, next} headers; a warp takes chunks from the pool with one atomic per chunk_slots slots;
//   sample   = {q.xyz, v.xyz, r, tag}  tag = zones | zone_index << 2 (never 0);  after media_kernel the first 16 bytes
//              are {e.r, e.g, e.b, s} with s = -1 for a sample that did not pass the 0.001 density gate (:71);
//   state    = {kState} header + 32 slots {p.xyz, v.xyz, steps | end << 28, 0}: exit state of a queued tile.
// Headers and state slots carry tag 0, which is how media_kernel tells them from samples.
// A pass ends when the pool passes its high-water mark (the warps stop taking tiles); a tile that cannot get a slot
// gives up, is put on the pass's redo list and traced again by the next pass; whatever is left after the last pass
// enqueued by the host is rendered by sweep_kernel (the fused code), so the frame is complete for any pool size.
#pragma once

namespace rrtk {
struct PassCtrl {
    unsigned cursor;          // slots handed out to warps so far (chunk granularity; may overshoot capacity)
    unsigned full;            // an allocation failed: stop taking tiles
    unsigned pend_count;      // tiles queued for the fold
    unsigned media_ticket, fold_ticket, redo_ticket;
    unsigned redo_out_count;  // tiles given up in this pass
    unsigned worked;          // tiles this pass took
};
struct PendTile {
    unsigned tile, first, end, state;
};
struct SplitArgs {
    uint4* slots;             // 2 x uint4 per slot
    unsigned capacity;        // slots
    unsigned high_water;      // stop taking tiles beyond this cursor
    unsigned chunk_slots;     // power of two
    unsigned chunk_shift;
    unsigned* chunk_used;     // slots in use per chunk (0 = chunk never closed: skipped by media_kernel)
    PassCtrl* pc;             // this pass
    const PassCtrl* pc_prev;  // previous pass (its redo list is this pass's first work), or null
    const unsigned* redo_in;
    unsigned* redo_out;
    unsigned redo_cap;
    PendTile* pend;
    unsigned* stats;          // [0] split passes that took tiles, [1] tiles rendered by sweep_kernel, [2] tiles taken by the split passes
    unsigned pass;
};
struct SplitKernels {
    void (*trace[2])(const FrameArgs, const SplitArgs);   // [spin != 0]
    void (*media)(const FrameArgs, const SplitArgs);
    void (*fold)(const FrameArgs, const SplitArgs);
    void (*sweep[2])(const FrameArgs, const SplitArgs);
};
const SplitKernels* rrt_split_kernels_strict();
const SplitKernels* rrt_split_kernels_fmad();
}  // namespace rrtk

using rrtk::PassCtrl;
using rrtk::PendTile;
using rrtk::SplitArgs;

namespace {

constexpr unsigned kNone = 0xffffffffu;   // "no slot" / allocation failure / "no tile"
enum : unsigned { kRec = 1u, kJump = 2u, kState = 3u };
constexpr int kMediaBlock = 128;          // media_kernel CTA
constexpr int kMediaBatch = 128;          // slots one warp of media_kernel takes per ticket

// pixel of a lane in a tile (the centre-outwards tile order of render_kernel)
struct TilePix {
    int x, y, ly;
    bool valid;
};
__device__ __forceinline__ TilePix tile_pixel(const FrameArgs& A, unsigned tile, int lane) {
    const int ntx = (A.w + kRTileW - 1) / kRTileW;
    const int nty = (A.local_rows + kRTileH - 1) / kRTileH;
    const int tx = (int)(tile % (unsigned)ntx), k = (int)(tile / (unsigned)ntx);
    const int c = nty >> 1, m = min(c, nty - 1 - c);
    int ty;
    if (k <= 2 * m) ty = (k & 1) ? c + ((k + 1) >> 1) : c - (k >> 1);
    else ty = (c > nty - 1 - c) ? (c - m - 1) - (k - (2 * m + 1)) : (c + m + 1) + (k - (2 * m + 1));
    TilePix t;
    t.x = tx * kRTileW + (lane & (kRTileW - 1));
    t.ly = ty * kRTileH + lane / kRTileW;
    const int grp = t.ly / A.band_group;
    t.y = (grp * A.band_nranks + A.band_rank) * A.band_group + (t.ly - grp * A.band_group);
    t.valid = t.x < A.w && t.ly < A.local_rows && t.y < A.h;
    return t;
}
__device__ __forceinline__ unsigned num_tiles(const FrameArgs& A) {
    return (unsigned)(((A.w + kRTileW - 1) / kRTileW) * ((A.local_rows + kRTileH - 1) / kRTileH));
}

// Next tile of a pass (lane 0 decides, everyone gets it): the previous pass's redo list first, then fresh tickets.
// `stop` makes a split pass end early (pool beyond its high-water mark).
__device__ __forceinline__ unsigned next_tile(const FrameArgs& A, const SplitArgs& S, unsigned ntiles, bool may_stop) {
    unsigned tile = kNone;
    if ((threadIdx.x & 31) == 0) {
        bool stop = false;
        if (may_stop) {
            const volatile PassCtrl* pc = S.pc;
            stop = pc->full != 0u || pc->cursor > S.high_water;
        }
        if (!stop) {
            const unsigned n_redo = S.pc_prev ? S.pc_prev->redo_out_count : 0u;
            if (n_redo) {
                const unsigned t = atomicAdd(&S.pc->redo_ticket, 1u);
                if (t < n_redo && t < S.redo_cap) tile = S.redo_in[t];
            }
            if (tile == kNone) {
                const unsigned t = atomicAdd(A.ticket, 1u);
                if (t < ntiles) tile = t;
            }
        }
    }
    return __shfl_sync(0xffffffffu, tile, 0);
}

// ---- emitter: the record stream of one tracing warp ---------------------------------------------------------------
// `state` (shared memory, one per warp) = end << 32 | cur: the warp's current chunk is [.., end] with `end` reserved for
// the jump header, cur the next free slot; 0 = no chunk yet.  Allocation is a CAS on that word, so it is correct
// whichever lanes of the warp happen to call it together (in practice: all the lanes that are inside a zone).
__device__ __forceinline__ void pool_put(uint4* slots, unsigned slot, uint4 a, uint4 b) {
    slots[2ull * slot] = a;
    slots[2ull * slot + 1] = b;
}
// The rare part of an allocation: the warp's chunk is used up (or there is none yet, or another group of the same warp got
// in between).  Takes a new chunk from the pool, chains it behind a jump header, retries.  Returns the first of `need`
// consecutive slots or kNone (pool exhausted).  Called by one lane; out of line, so the tracing loop keeps nothing of it live.
__device__ __noinline__ unsigned pool_alloc_slow(const SplitArgs* S, unsigned long long* state, unsigned need) {
    for (;;) {
        const unsigned long long st = *(volatile unsigned long long*)state;
        const unsigned cur = (unsigned)st, end = (unsigned)(st >> 32);
        if (end != 0u && cur + need <= end) {
            if (atomicCAS(state, st, ((unsigned long long)end << 32) | (cur + need)) != st) continue;
            return cur;
        }
        if (*(volatile unsigned*)&S->pc->full) return kNone;
        const unsigned nb = atomicAdd(&S->pc->cursor, S->chunk_slots);
        if (nb > S->capacity || S->capacity - nb < S->chunk_slots) {
            atomicExch(&S->pc->full, 1u);
            return kNone;
        }
        const unsigned nend = nb + S->chunk_slots - 1u;
        if (atomicCAS(state, st, ((unsigned long long)nend << 32) | (nb + need)) != st) continue;   // (chunk nb is lost: stays unused)
        if (end != 0u) {   // close the old chunk behind a jump header
            pool_put(S->slots, cur, make_uint4(kJump, nb, 0u, 0u), make_uint4(0u, 0u, 0u, 0u));
            S->chunk_used[cur >> S->chunk_shift] = (cur & (S->chunk_slots - 1u)) + 1u;
        }
        return nb;
    }
}
// `need` consecutive slots, the first of them a header {kind, w1}; returns the header's slot or kNone.  One lane calls it on
// behalf of the lanes that emit together.  The common case -- room in the warp's chunk, nobody in between -- is one shared-
// memory load and one CAS, inline (a disk-plane ray emits at every one of its 2000 steps: this is on the frame's longest
// dependency chain).
__device__ __forceinline__ unsigned pool_alloc(const SplitArgs& S, unsigned long long* state, unsigned need, unsigned kind, unsigned w1) {
    const unsigned long long st = *(volatile unsigned long long*)state;
    const unsigned cur = (unsigned)st, end = (unsigned)(st >> 32);
    unsigned base;
    if (end != 0u && cur + need <= end && atomicCAS(state, st, ((unsigned long long)end << 32) | (cur + need)) == st) base = cur;
    else base = pool_alloc_slow(&S, state, need);
    if (base != kNone) pool_put(S.slots, base, make_uint4(kind, w1, 0u, 0u), make_uint4(0u, 0u, 0u, 0u));
    return base;
}

struct Emitter {
    const SplitArgs& S;
    unsigned long long* state;
    unsigned first;     // first record this lane wrote for the current tile
    bool gave_up;       // the pool had no room: the tile goes to the redo list
    float isco, disk_out;

    __device__ __forceinline__ Emitter(const SplitArgs& s, unsigned long long* st, const Consts& C)
        : S(s), state(st), first(kNone), gave_up(false), isco(C.isco), disk_out(C.disk_out) {}

    __device__ __forceinline__ void put(unsigned slot, uint4 a, uint4 b) const { pool_put(S.slots, slot, a, b); }
    __device__ __forceinline__ unsigned alloc(unsigned need, unsigned kind, unsigned w1) const { return pool_alloc(S, state, need, kind, w1); }
    // one in-zone sample (trace_ray, kMediaEmit)
    __device__ __forceinline__ void emit(V3 q, V3 v, float r, int zone_index, unsigned z) {
        if (gave_up) return;
        // Outside the ring ISCO <= R <= DISK_OUT both density functions return 0 before anything else
        // (densities.h:21-22, 70-71): such a sample cannot pass the gate of raymarcher.cu:71 and is not stored.
        const float R = sqrtf(rrt::ring_r2(q));
        if (R < isco || R > disk_out) return;
#if defined(RRT_DBG_EMIT) && RRT_DBG_EMIT == 1   // timing experiments only (wrong frames): where does an emission's time go?
        return;
#endif
        const unsigned grp = __activemask();
        const unsigned lane = threadIdx.x & 31u;
        const unsigned leader = (unsigned)__ffs((int)grp) - 1u, n = (unsigned)__popc(grp);
        const unsigned rank = (unsigned)__popc(grp & ((1u << lane) - 1u));
        unsigned base = 0u;
#if defined(RRT_DBG_EMIT) && RRT_DBG_EMIT == 3
        base = (blockIdx.x & 1023u) * 64u;
#else
        if (lane == leader) base = alloc(n + 1u, kRec, grp);
        base = __shfl_sync(grp, base, (int)leader);
        if (base == kNone) { gave_up = true; return; }
        first = min(first, base);
#endif
#if defined(RRT_DBG_EMIT) && RRT_DBG_EMIT == 2
        return;
#endif
        put(base + 1u + rank, make_uint4(__float_as_uint(q.x), __float_as_uint(q.y), __float_as_uint(q.z), __float_as_uint(v.x)),
            make_uint4(__float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(r), z | ((unsigned)zone_index << 2)));
    }
    __device__ __forceinline__ unsigned cursor() const { return (unsigned)*(volatile unsigned long long*)state; }
    // end of the kernel: the warp's last chunk becomes visible to media_kernel
    __device__ __forceinline__ void close() const {
        const unsigned long long st = *(volatile unsigned long long*)state;
        const unsigned cur = (unsigned)st, end = (unsigned)(st >> 32);
        if (end != 0u) S.chunk_used[end >> S.chunk_shift] = cur & (S.chunk_slots - 1u);
    }
};

struct TileCounters {
    unsigned long long steps = 0, disk = 0, dust = 0, dense = 0;
    unsigned cap = 0, esc = 0, exh = 0, touch = 0;
    __device__ __forceinline__ void flush(unsigned long long* counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            steps += __shfl_xor_sync(0xffffffffu, steps, o);
            disk += __shfl_xor_sync(0xffffffffu, disk, o);
            dust += __shfl_xor_sync(0xffffffffu, dust, o);
            dense += __shfl_xor_sync(0xffffffffu, dense, o);
            cap += __shfl_xor_sync(0xffffffffu, cap, o);
            esc += __shfl_xor_sync(0xffffffffu, esc, o);
            exh += __shfl_xor_sync(0xffffffffu, exh, o);
            touch += __shfl_xor_sync(0xffffffffu, touch, o);
        }
        if ((threadIdx.x & 31) == 0 && counters) {
            if (steps) atomicAdd(counters + 0, steps);
            if (disk) atomicAdd(counters + 1, disk);
            if (dust) atomicAdd(counters + 2, dust);
            if (dense) atomicAdd(counters + 3, dense);
            if (cap) atomicAdd(counters + 4, (unsigned long long)cap);
            if (esc) atomicAdd(counters + 5, (unsigned long long)esc);
            if (exh) atomicAdd(counters + 6, (unsigned long long)exh);
            if (touch) atomicAdd(counters + 7, (unsigned long long)touch);
        }
    }
};

// ---- pass kernel 1: trajectories ------------------------------------------------------------------------------------
template <bool SPIN>
__global__ void __launch_bounds__(kRenderBlock, RRT_MIN_BLOCKS) trace_kernel(const __grid_constant__ FrameArgs A,
                                                                              const __grid_constant__ SplitArgs S) {
    __shared__ unsigned long long em_state[kRenderBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) em_state[warp] = 0ull;
    __syncwarp();
    const unsigned ntiles = num_tiles(A);
    TileCounters cnt;
    Emitter em(S, &em_state[warp], A.C);
    unsigned taken = 0;

    for (;;) {
        const unsigned tile = next_tile(A, S, ntiles, true);
        if (tile == kNone) break;
        ++taken;
        const TilePix px = tile_pixel(A, tile, lane);
        RayResult R;
        R.steps = 0; R.n_disk = R.n_dust = R.n_dense = 0;
        R.captured = R.touched = R.exhausted = false;
        R.p = R.v = mk(0.f, 0.f, 0.f);
        R.T = 1.0f; R.I[0] = R.I[1] = R.I[2] = 0.f;
        R.uvx = R.uvy = 0.f;
        em.first = kNone;
        em.gave_up = false;
#ifdef RRT_WITH_TILE_LOG
        unsigned long long t_begin = 0;
        if (A.tile_log) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
#ifdef RRT_WITH_GROUP_STEPS
        __syncwarp();
        if (lane == 0) g_dbg_group_steps[warp] = 0u;
        __syncwarp();
#endif
#endif
        if (px.valid) trace_ray<SPIN, kMediaEmit>(A, px.x, px.y, R, em);
        __syncwarp();
#ifdef RRT_WITH_TILE_LOG   // profiling build only (rrt_debug_tile_log): when did this warp trace which tile
        if (A.tile_log) {
            const int most = __reduce_max_sync(0xffffffffu, R.steps);
            if (lane == 0 && tile < A.tile_log_cap) {
                unsigned long long t_end;
                unsigned smid;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                unsigned long long* e = A.tile_log + 4ull * tile;
                const int ntx = (A.w + kRTileW - 1) / kRTileW;
                e[0] = t_begin; e[1] = t_end; e[2] = ((unsigned long long)(px.ly / kRTileH) << 32) | (unsigned)(tile % (unsigned)ntx);
                e[3] = ((unsigned long long)smid << 32) | (unsigned)most | (dbg_group_steps(warp) << 12);
            }
        }
#endif
        const unsigned first = __reduce_min_sync(0xffffffffu, em.first);
        bool gave_up = __any_sync(0xffffffffu, em.gave_up);
        unsigned end_pos = 0u, state_pos = 0u;
        if (!gave_up && first != kNone) {   // the tile has samples: its exit states go to the pool, the fold finishes it
            end_pos = em.cursor();
            if (lane == 0) state_pos = em.alloc(33u, kState, tile);
            state_pos = __shfl_sync(0xffffffffu, state_pos, 0);
            gave_up = state_pos == kNone;
        }
        if (gave_up) {   // no room in the pool: the next pass (or the sweep) traces this tile again
            if (lane == 0) {
                const unsigned i = atomicAdd(&S.pc->redo_out_count, 1u);
                if (i < S.redo_cap) S.redo_out[i] = tile;
            }
            continue;
        }
        const unsigned end = (R.captured ? kEndCaptured : 0u) | (R.exhausted ? kEndExhausted : 0u);
        if (first == kNone) {
            if (px.valid) {
                if (R.captured) R.T = 0.0f;
                finish_ray_inl(A, px.x, px.y, px.ly, R.uvx, R.uvy, 0.f, 0.f, 0.f, R.T, R.p, R.v, R.steps, end);
            }
        } else {
            em.put(state_pos + 1u + (unsigned)lane,
                   make_uint4(__float_as_uint(R.p.x), __float_as_uint(R.p.y), __float_as_uint(R.p.z), __float_as_uint(R.v.x)),
                   make_uint4(__float_as_uint(R.v.y), __float_as_uint(R.v.z), (unsigned)R.steps | (end << 28), 0u));
            if (lane == 0) {
                const unsigned i = atomicAdd(&S.pc->pend_count, 1u);
                S.pend[i] = PendTile{tile, first, end_pos, state_pos};
            }
        }
        if (px.valid) {
            cnt.steps += (unsigned)R.steps;
            cnt.disk += R.n_disk; cnt.dust += R.n_dust;
            cnt.cap += R.captured; cnt.exh += R.exhausted; cnt.esc += (!R.captured && !R.exhausted);
        }
    }
    __syncwarp();
    if (lane == 0) {
        em.close();
        if (taken) {
            atomicAdd(&S.pc->worked, taken);
            atomicMax(S.stats + 0, S.pass + 1u);
            atomicAdd(S.stats + 2, taken);
        }
    }
    cnt.flush(A.counters);
}

// ---- pass kernel 2: the samples, one thread per slot ----------------------------------------------------------------
// A warp takes kMediaBatch consecutive slots per ticket.  Stage 1 evaluates, lane per slot, the disk density and the
// dust envelope (both cheap to moderately expensive, and nearly every lane has the same work: consecutive slots are
// the lanes of one tracing warp at one step).  The dust strands -- 19 value-noise evaluations, needed only where the
// envelope survived its 0.001 cut (densities.h:84) -- are collected over the whole batch and evaluated 32 at a time.
// Stage 3 is media_final for every sample.
__global__ void __launch_bounds__(kMediaBlock, 4) media_kernel(const __grid_constant__ FrameArgs A, const __grid_constant__ SplitArgs S) {
    __shared__ float s_dd[kMediaBlock / 32][kMediaBatch];
    __shared__ float s_dc[kMediaBlock / 32][kMediaBatch];      // dust envelope, then dust density
    __shared__ unsigned short s_list[kMediaBlock / 32][kMediaBatch];
    const Consts& C = A.C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned n_slots = min(S.pc->cursor, S.capacity);
    float* dd_w = s_dd[warp];
    float* dc_w = s_dc[warp];
    unsigned short* list = s_list[warp];
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(&S.pc->media_ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        const unsigned long long base64 = (unsigned long long)t * kMediaBatch;
        if (base64 >= n_slots) break;
        const unsigned base = (unsigned)base64;
        unsigned n_list = 0;
        // stage 1
#pragma unroll 1
        for (int k = 0; k < kMediaBatch / 32; ++k) {
            const unsigned i = (unsigned)k * 32u + (unsigned)lane, s = base + i;
            unsigned tag = 0u;
            uint4 a = make_uint4(0u, 0u, 0u, 0u);
            if (s < n_slots && (s & (S.chunk_slots - 1u)) < S.chunk_used[s >> S.chunk_shift]) {
                tag = S.slots[2ull * s + 1].w;
                if (tag & 3u) a = S.slots[2ull * s];
            }
            float dd = 0.0f, base_d = 0.0f;
            const V3 q = mk(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
            if (tag & 1u) dd = rrt::disk_density(C, q, A.time);                               // :68
            if (tag & 2u) base_d = rrt::dust_base(C, q);                                      // :69, densities.h:70-84
            dd_w[i] = dd;
            dc_w[i] = 0.0f;
            const bool strands = base_d != 0.0f;
            const unsigned m = __ballot_sync(0xffffffffu, strands);
            if (strands) {
                list[n_list + (unsigned)__popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
                dc_w[i] = base_d;
            }
            n_list += (unsigned)__popc(m);
        }
        __syncwarp();
        // stage 2: dust strands of the survivors, 32 at a time
#pragma unroll 1
        for (unsigned j0 = 0; j0 < n_list; j0 += 32u) {
            const unsigned j = j0 + (unsigned)lane;
            if (j < n_list) {
                const unsigned i = list[j];
                const uint4 a = S.slots[2ull * (base + i)];
                const V3 q = mk(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
                dc_w[i] = rrt::dust_strands(C, q, A.time, dc_w[i]);
            }
        }
        __syncwarp();
        // stage 3: emission colour and step transmittance
#pragma unroll 1
        for (int k = 0; k < kMediaBatch / 32; ++k) {
            const unsigned i = (unsigned)k * 32u + (unsigned)lane, s = base + i;
            unsigned tag = 0u;
            if (s < n_slots && (s & (S.chunk_slots - 1u)) < S.chunk_used[s >> S.chunk_shift]) tag = S.slots[2ull * s + 1].w;
            if (tag & 3u) {
                const uint4 a = S.slots[2ull * s], b = S.slots[2ull * s + 1];
                const V3 q = mk(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
                const V3 v = mk(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
                const MediaOut m = media_final(C, q, v, __uint_as_float(b.z), C.h[(tag >> 2) & 3u], dd_w[i], dc_w[i]);
                S.slots[2ull * s] = make_uint4(__float_as_uint(m.er), __float_as_uint(m.eg), __float_as_uint(m.eb),
                                               __float_as_uint(m.dense ? m.s : -1.0f));
            }
        }
        __syncwarp();
    }
}

// ---- pass kernel 3: fold + finish, one warp per queued tile ----------------------------------------------------------
__global__ void __launch_bounds__(128) fold_kernel(const __grid_constant__ FrameArgs A, const __grid_constant__ SplitArgs S) {
    const int lane = threadIdx.x & 31;
    const unsigned n_pend = S.pc->pend_count;
    const unsigned lt = (1u << lane) - 1u;
    TileCounters cnt;
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(&S.pc->fold_ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_pend) break;
        const PendTile e = S.pend[t];
        const TilePix px = tile_pixel(A, e.tile, lane);
        const uint4 sa = S.slots[2ull * (e.state + 1u + (unsigned)lane)], sb = S.slots[2ull * (e.state + 1u + (unsigned)lane) + 1];
        const V3 p = mk(__uint_as_float(sa.x), __uint_as_float(sa.y), __uint_as_float(sa.z));
        const V3 v = mk(__uint_as_float(sa.w), __uint_as_float(sb.x), __uint_as_float(sb.y));
        const int steps = (int)(sb.z & 0x0fffffffu);
        unsigned end = sb.z >> 28;
        float Ir = 0.f, Ig = 0.f, Ib = 0.f, T = 1.0f;
        unsigned n_dense = 0;
        unsigned pos = e.first;
        // The walk is a chain of dependent loads (a record's length is in its header).  Each round trip therefore brings
        // TWO windows: the record at `pos` (header + up to 32 samples: lane l loads slot pos + l, lane 31 also slot
        // pos + 32) and the 33 slots behind a guessed next header -- the same length as the previous record, which is what
        // a tile whose lanes stay inside a zone produces step after step.  A right guess folds two records per trip.
        // (the bound only keeps a corrupt stream from spinning: a tile has at most 32 records per step and a jump per record)
        unsigned guess = 33u;
        for (unsigned guard = ((unsigned)A.C.max_steps + 2u) * 64u; pos != e.end && guard; --guard) {
            const unsigned pos2 = pos + guess;
            uint4 wv[2], w32[2];
            wv[0] = S.slots[2ull * (pos + (unsigned)lane)];
            wv[1] = pos2 + 33u <= S.capacity + 64u ? S.slots[2ull * (pos2 + (unsigned)lane)] : make_uint4(0u, 0u, 0u, 0u);
            w32[0] = w32[1] = make_uint4(0u, 0u, 0u, 0u);
            if (lane == 31) {
                w32[0] = S.slots[2ull * (pos + 32u)];
                if (pos2 + 33u <= S.capacity + 64u) w32[1] = S.slots[2ull * (pos2 + 32u)];
            }
            bool stop = false;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k == 1 && (pos != pos2 || pos == e.end)) break;   // the guess was wrong (or the tile ends here): next trip
                const unsigned kind = __shfl_sync(0xffffffffu, wv[k].x, 0), w1 = __shfl_sync(0xffffffffu, wv[k].y, 0);
                if (kind == kJump) { pos = w1; break; }
                if (kind != kRec) { stop = true; break; }   // (corrupt stream: cannot happen; do not spin)
                const unsigned mask = w1, idx = 1u + (unsigned)__popc(mask & lt);
                const int src = (int)(idx & 31u);
                float er = __uint_as_float(__shfl_sync(0xffffffffu, wv[k].x, src)), eg = __uint_as_float(__shfl_sync(0xffffffffu, wv[k].y, src));
                float eb = __uint_as_float(__shfl_sync(0xffffffffu, wv[k].z, src)), sx = __uint_as_float(__shfl_sync(0xffffffffu, wv[k].w, src));
                const float er32 = __uint_as_float(__shfl_sync(0xffffffffu, w32[k].x, 31)), eg32 = __uint_as_float(__shfl_sync(0xffffffffu, w32[k].y, 31));
                const float eb32 = __uint_as_float(__shfl_sync(0xffffffffu, w32[k].z, 31)), s32 = __uint_as_float(__shfl_sync(0xffffffffu, w32[k].w, 31));
                if (idx == 32u) { er = er32; eg = eg32; eb = eb32; sx = s32; }
                if (((mask >> lane) & 1u) && sx != -1.0f) {                                   // :71
                    ++n_dense;
                    const float wgt = rrt::mul(rrt::sub(1.0f, sx), T);                        // :109
                    Ir = rrt::mad(er, wgt, Ir); Ig = rrt::mad(eg, wgt, Ig); Ib = rrt::mad(eb, wgt, Ib);   // :111-113
                    T = rrt::mul(T, sx);                                                      // :115
                }
                guess = 1u + (unsigned)__popc(mask);
                pos += guess;
            }
            if (stop) break;
        }
        if (n_dense) end |= kEndTouched;
        if (end & kEndCaptured) T = 0.0f;                                                     // :49
        if (px.valid) {
            float uvx, uvy;
            pixel_uv(A, px.x, px.y, uvx, uvy);
            finish_ray_inl(A, px.x, px.y, px.ly, uvx, uvy, Ir, Ig, Ib, T, p, v, steps, end);
            cnt.dense += n_dense;
            cnt.touch += n_dense ? 1u : 0u;
        }
    }
    cnt.flush(A.counters);
}

// ---- after the last pass: whatever the split passes left, with the fused code ------------------------------------------
template <bool SPIN>
__global__ void __launch_bounds__(kRenderBlock, RRT_MIN_BLOCKS_MEDIA) sweep_kernel(const __grid_constant__ FrameArgs A,
                                                                                    const __grid_constant__ SplitArgs S) {
    const int lane = threadIdx.x & 31;
    const unsigned ntiles = num_tiles(A);
    TileCounters cnt;
    unsigned taken = 0;
    for (;;) {
        const unsigned tile = next_tile(A, S, ntiles, false);
        if (tile == kNone) break;
        ++taken;
        const TilePix px = tile_pixel(A, tile, lane);
        if (!px.valid) continue;
        RayResult R;
        NoEmit no_emit;
        trace_ray<SPIN, kMediaInline>(A, px.x, px.y, R, no_emit);
        finish_ray_inl(A, px.x, px.y, px.ly, R.uvx, R.uvy, R.I[0], R.I[1], R.I[2], R.T, R.p, R.v, R.steps,
                       (R.captured ? kEndCaptured : 0u) | (R.touched ? kEndTouched : 0u) | (R.exhausted ? kEndExhausted : 0u));
        cnt.steps += (unsigned)R.steps;
        cnt.disk += R.n_disk; cnt.dust += R.n_dust; cnt.dense += R.n_dense;
        cnt.cap += R.captured; cnt.exh += R.exhausted; cnt.esc += (!R.captured && !R.exhausted); cnt.touch += R.touched;
    }
    if (lane == 0 && taken) atomicAdd(S.stats + 1, taken);
    if (blockIdx.x == 0 && threadIdx.x == 0) { S.stats[3] = S.pass; S.stats[4] = ntiles; }   // what the host enqueued, for its next guess
    cnt.flush(A.counters);
}

const rrtk::SplitKernels kSplitKernels = {
    {trace_kernel<false>, trace_kernel<true>}, media_kernel, fold_kernel, {sweep_kernel<false>, sweep_kernel<true>},
};
}  // namespace

namespace rrtk {
#if RRT_FMAD
const SplitKernels* rrt_split_kernels_fmad() { return &kSplitKernels; }
#else
const SplitKernels* rrt_split_kernels_strict() { return &kSplitKernels; }
#endif
}  // namespace rrtk
