/* rr_kernels.cu — sm_100a kernels of the RadaRays hot path (RadarCPU::simulate, RadarCPU.cpp:30-564).
 *
 * rr_trace_kernel, once per pass (persistent, barrier-free, one WARP per group of 32 waves of the launch-wide list):
 *   beam bundle generation      (RadarCPU.cpp:184-209, radar_algorithms.cpp:150-169)
 *   closest-hit traversal       (Rmagine/Embree call at RadarCPU.cpp:236) over the 32-byte quantised BVH
 *   move + Snell/Fresnel + BRDF (radar_types.h:108-120, radar_algorithms.h:55-139,168-187, RadarCPU.cpp:243-371)
 *   pruning + ordered respawn   (RadarCPU.cpp:288-290,364-389) via ballot/popc compaction (keeps the reference's list order)
 * rr_scan_kernel between passes: prefix sums that turn per-group child counts into the next pass's list (no data moved)
 * rr_draw_kernel (one CTA per (pose, azimuth)):
 *   range-bin accumulation      (RadarCPU.cpp:402-450) into a shared-memory column, in reference order, no atomics
 *   energy_max / ambient noise / normalise / mono8 (RadarCPU.cpp:453-542)
 * The arithmetic is written independently of oracle/rr_oracle.cpp; only the elementary primitives of
 * rr_detmath.h are shared. Compile with --fmad=false: every fused multiply-add below is explicit.
 */
#include <algorithm>
#include <cooperative_groups.h>
#include "rr_internal.h"

#define RR_FULL 0xffffffffu
#ifndef RR_MIN_BLOCKS
#define RR_MIN_BLOCKS 10           /* resident 128-thread trace CTAs per SM the register allocation is tuned for */
#endif

/* Out-of-line helpers of the shading code. The shading of a wave is ~4000 straight-line instructions executed once per
 * ray, far more than the instruction caches hold next to the walk loop (20 % of the trace kernel's stall samples were
 * "no instruction"); sharing the quaternion rotations, normalisations and IEEE double divisions as real functions
 * instead of inlining every use shrinks that footprint (measured: 1.716 -> 1.67 ms per 16-pose step). Same arithmetic. */
__device__ __noinline__ rr_vec3 rr_qrot_ol(rr_quat q, rr_vec3 v) { return rr_qrot(q, v); }
__device__ __noinline__ double rr_ddiv_ol(double a, double b) { return a / b; }
__device__ __noinline__ rr_vec3 rr_normalize_ol(rr_vec3 v) { return rr_normalize(v); }
#define RR_QROT(q, v) rr_qrot_ol(q, v)
#define RR_DDIV(a, b) rr_ddiv_ol(a, b)
#define RR_NORMALIZE(v) rr_normalize_ol(v)
/* rr_tan_of (rr_detmath.h:111) with its one division out of line: (-c)/s or s/c */
__device__ __forceinline__ double rr_tan_of_ol(rr_sincos_t p) { return RR_DDIV((p.q & 1) ? -p.c : p.s, (p.q & 1) ? p.s : p.c); }

/* (float)acos(0.0f) as the reference's acos(float) gives it (radar_algorithms.h:106 with a zero refraction vector) */
#define RR_ACOSF_OF_ZERO 1.57079637050628662109375f

/* back_reflection_shader (radar_algorithms.h:168-187) without the incoming energy: A + B * pow(cos(angle), C) with
 * (A, B, C) = (ambient, diffuse, specular). With a lobe factor B = +-0 (config/mulran_kaist02.yaml:15-18 and most of the
 * reference's material files) the lobe only contributes the SIGN of a zero as long as it is finite: for angle < pi/2 the
 * cosine is a float in (0, 1] (no float equals pi/2, so it never rounds to 0), and pow of that with an exponent C >= 0 is
 * a finite value >= +0, hence B * lobe = B (as +-0) exactly. Everything else takes the full evaluation. */
__device__ __forceinline__ float rr_brdf(float angle, float4 mt)
{
    float term;
    if (mt.z == 0.0f && mt.w >= 0.0f && (double)angle < RR_PIO2_HI) term = mt.z;
    else term = mt.z * rr_powf(rr_cosf(angle), mt.w);
    return mt.y * 1.0f + term;
}

/* Wave state. Quirk kept on purpose: the reference never updates DirectedWave::velocity on the waves it pushes
 * (RadarCPU.cpp:285-286,364-365 copy only dir and energy out of fresnel()'s result), so every wave travels with
 * the initial 0.3 m/ns (RadarCPU.cpp:110) and only material_id tracks the medium. velocity is therefore a
 * constant here, not per-wave state. */
#define RR_WAVE_VELOCITY 0.3
struct RRWave {
    rr_vec3 o, d;
    double energy, time;
    uint32_t mat;
};

/* ------------------------------------------------------------------------------------------------
 * closest hit: smallest t, ties -> lowest face id (rr_detmath.h). Returns triangle SLOT or -1.
 * ---------------------------------------------------------------------------------------------- */
/* reciprocal that stays finite for zero components: the slab planes then sit at -/+ 1e30-ish, on the same side when
 * the origin is outside the slab (reject), on opposite sides when inside (accept) — conservative, and no 0*inf NaNs */
__device__ __forceinline__ float rr_safe_rcp(float d) { return 1.0f / ((fabsf(d) > 1e-30f) ? d : copysignf(1e-30f, d)); }

/* one PRMT: picks the low or high u16 of `w` (selector 0x7610 / 0x7632) under the exponent bytes of 2^23,
 * i.e. returns the float 8388608 + q without any int->float conversion */
__device__ __forceinline__ float rr_plane(uint32_t w, uint32_t sel, uint32_t magic)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "r"(sel));
    return __uint_as_float(r);
}

/* Blackwell's packed single-precision FMA (fma.rn.f32x2 -> FFMA2): two independent IEEE fmas per issue slot. The walk
 * is issue-bound, so pairing the twelve plane distances of a node step (x with y of one child and plane kind, the z
 * planes of both children) halves their FFMA issue slots. Element results are those of two scalar fmaf. */
#ifndef RR_FFMA2
#define RR_FFMA2 0
#endif
__device__ __forceinline__ uint64_t rr_pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void rr_ffma2(float& x, float& y, float fx, float fy, uint64_t a, uint64_t b)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(rr_pack2(fx, fy)), "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(d));
}

__device__ __forceinline__ uint32_t rr_pop_local(uint32_t& sp)
{
    uint32_t v;
    sp -= 4;
    asm volatile("{ .reg .u64 a; cvt.u64.u32 a, %1; ld.local.u32 %0, [a]; }" : "=r"(v) : "r"(sp) : "memory");
    return v;
}

/* Traversal stack: the first RR_SMEM_STACK entries of every thread live in shared memory, laid out [entry][thread]
 * (a warp's accesses never conflict whatever the lanes' depths are); only deeper entries (rare: the stack holds one
 * postponed sibling per "both children hit" node of the current path) go to the thread's local-memory array. */
#ifndef RR_SMEM_STACK
#define RR_SMEM_STACK 0             /* measured on the B200 (urban-5M, 16 poses): 0 -> 1.591 ms, 16 -> 1.697, 24 -> 1.695, 32 -> 1.735: the shared-memory stack costs L1 capacity (the node gathers live there) and a branch per push/pop */
#endif
#ifndef RR_STACK_PTR
#define RR_STACK_PTR 1
#endif
/* RR_EARLY_EXIT = T > 0: the warp leaves the node loop as soon as at most T of its lanes are still descending (one vote
 * per node step) instead of waiting for the last one; the stragglers idle through the leaf tests and walk on afterwards.
 * Every ray still visits its own nodes in its own order, so hits and counters do not change. 0 = plain while-while. */
#ifndef RR_LDG256
#define RR_LDG256 0
#endif
#ifndef RR_EARLY_EXIT
#define RR_EARLY_EXIT 0
#endif
#define RR_TRACE_SMEM_BYTES ((size_t)RR_SMEM_STACK * RR_TRACE_BLOCK * sizeof(uint32_t))

template <bool STATS>
__device__ __forceinline__ int rr_trace(const RRNode* __restrict__ nodes, const float4* __restrict__ tris,
                                        uint32_t root_ref, const float* go, const float* gs,
                                        rr_vec3 o, rr_vec3 d, float tmax, float& t_hit, int& face_hit,
                                        unsigned& n_nodes, unsigned& n_tris, uint32_t* s_stack, unsigned wmask)
{
    /* plane distance in ray space (rr_internal.h): t = f*A + B', f = 2^23 + q. One PRMT + one FFMA per plane; the
     * near/far plane of each axis is chosen by the per-ray selectors, so no per-axis min/max is needed. */
    const float ix = rr_safe_rcp(d.x), iy = rr_safe_rcp(d.y), iz = rr_safe_rcp(d.z);
    const float sax = gs[0] * ix, say = gs[1] * iy, saz = gs[2] * iz;
    const float sbx = fmaf(-8388608.0f, sax, (go[0] - o.x) * ix);
    const float sby = fmaf(-8388608.0f, say, (go[1] - o.y) * iy);
    const float sbz = fmaf(-8388608.0f, saz, (go[2] - o.z) * iz);
    const uint32_t nx = (ix >= 0.f) ? 0x7610u : 0x7632u, fx = nx ^ 0x0022u;
    const uint32_t ny = (iy >= 0.f) ? 0x7610u : 0x7632u, fy = ny ^ 0x0022u;
    const uint32_t nz = (iz >= 0.f) ? 0x7610u : 0x7632u, fz = nz ^ 0x0022u;
    uint32_t magic;
    asm volatile("mov.b32 %0, 0x4B000000;" : "=r"(magic));   /* kept in a register: PRMT's third operand */
#if RR_FFMA2
    const uint64_t Axy = rr_pack2(sax, say), Bxy = rr_pack2(sbx, sby), Azz = rr_pack2(saz, saz), Bzz = rr_pack2(sbz, sbz);
#endif
#if RR_SMEM_STACK > 0
    uint32_t stack[RR_STACK_SIZE + 1 - RR_SMEM_STACK];
    uint32_t* const sst = s_stack + threadIdx.x;                     /* entry k of this thread: sst[k * blockDim.x] */
#define RR_PUSH(v) do { if (sp < RR_SMEM_STACK) sst[sp * RR_TRACE_BLOCK] = (v); else stack[sp - RR_SMEM_STACK] = (v); sp++; } while (0)
#define RR_POP() ((--sp < RR_SMEM_STACK) ? sst[sp * RR_TRACE_BLOCK] : stack[sp - RR_SMEM_STACK])
    int sp = 0;
#elif RR_STACK_PTR
    /* the stack pointer IS the entry's local-memory address: push and pop need no index scaling (one instruction less
     * each than stack[sp++] / stack[--sp]) */
    uint32_t stack[RR_STACK_SIZE + 1];
    uint32_t sp;                                       /* local-window addresses fit 32 bits */
    { uint64_t a; asm volatile("cvta.to.local.u64 %0, %1;" : "=l"(a) : "l"(stack) : "memory"); sp = (uint32_t)a; }
#define RR_PUSH(v) do { asm volatile("{ .reg .u64 a; cvt.u64.u32 a, %0; st.local.u32 [a], %1; }" :: "r"(sp), "r"(v) : "memory"); sp += 4; } while (0)
#define RR_POP() rr_pop_local(sp)
#else
    uint32_t stack[RR_STACK_SIZE + 1];
#define RR_PUSH(v) do { stack[sp++] = (v); } while (0)
#define RR_POP() (stack[--sp])
    int sp = 0;
#endif
    /* sentinel at the bottom: popping it ends the walk, so no pop ever tests for an empty stack */
    RR_PUSH(RR_REF_EMPTY);
    uint32_t cur = root_ref;
    float best_t = INFINITY;
    int best_face = -1, best_slot = -1;
    float limit = tmax * 1.00001f + 1e-6f;        /* prune slack >> rounding error of t (ties must be visited) */

    /* while-while walk: every lane first descends through inner nodes until it holds a leaf (or has nothing left); the
     * lanes of the warp reconverge after the inner loop, so the triangle tests below run once per round with (nearly)
     * all lanes on a leaf instead of being replayed for one or two lanes between node steps. RR_REF_EMPTY (which has
     * the leaf bit set, so it also ends the inner loop) marks a finished lane. */
#if RR_WIDE_BVH
    /* one step over a 4-wide node: four boxes, the hit children sorted by entry distance (keys of missed children are
     * +inf), next = nearest, the others postponed far to near */
    auto node_step = [&]() {
        const uint4* np = reinterpret_cast<const uint4*>(reinterpret_cast<const RRNode4*>(nodes) + cur);
        const uint4 A = __ldg(np), B = __ldg(np + 1), C = __ldg(np + 2);     /* c0.xyz c1.x | c1.yz c2.xy | c2.z c3.xyz */
        const uint4 R = __ldg(np + 3);
        if (STATS) n_nodes++;
#define RR_BOX(wx, wy, wz, key) do { \
            const float tn_ = fmaxf(fmaxf(fmaf(rr_plane(wx, nx, magic), sax, sbx), fmaf(rr_plane(wy, ny, magic), say, sby)), \
                                    fmaxf(fmaf(rr_plane(wz, nz, magic), saz, sbz), 0.0f)); \
            const float tf_ = fminf(fminf(fmaf(rr_plane(wx, fx, magic), sax, sbx), fmaf(rr_plane(wy, fy, magic), say, sby)), \
                                    fminf(fmaf(rr_plane(wz, fz, magic), saz, sbz), limit)); \
            key = (tn_ <= tf_) ? tn_ : INFINITY; } while (0)
        float k0, k1, k2, k3;
        RR_BOX(A.x, A.y, A.z, k0);
        RR_BOX(A.w, B.x, B.y, k1);
        RR_BOX(B.z, B.w, C.x, k2);
        RR_BOX(C.y, C.z, C.w, k3);
#undef RR_BOX
        uint32_t r0 = R.x, r1 = R.y, r2 = R.z, r3 = R.w;
#define RR_CE(ka, ra, kb, rb) do { const bool sw_ = kb < ka; const float kt_ = sw_ ? kb : ka; kb = sw_ ? ka : kb; ka = kt_; \
                                   const uint32_t rt_ = sw_ ? rb : ra; rb = sw_ ? ra : rb; ra = rt_; } while (0)
        RR_CE(k0, r0, k1, r1); RR_CE(k2, r2, k3, r3); RR_CE(k0, r0, k2, r2); RR_CE(k1, r1, k3, r3); RR_CE(k1, r1, k2, r2);
#undef RR_CE
        if (k0 < INFINITY) {
            if (k3 < INFINITY) RR_PUSH(r3);
            if (k2 < INFINITY) RR_PUSH(r2);
            if (k1 < INFINITY) RR_PUSH(r1);
            cur = r0;
        } else {
            cur = RR_POP();
        }
    };
#else
    /* one node step of this lane's ray: both child boxes, next = nearer hit child, the other one postponed */
    auto node_step = [&]() {
#if RR_LDG256
        uint4 a, b;                          /* the whole 32-byte node in one 256-bit load (sm_100: LDG.E.ENL2.256) */
        asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
            : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(nodes + cur));
#else
        const uint4* np = reinterpret_cast<const uint4*>(nodes + cur);
        const uint4 a = __ldg(np);           /* c0.x c0.y c0.z c1.x */
        const uint4 b = __ldg(np + 1);       /* c1.y c1.z ref0 ref1 */
#endif
        if (STATS) n_nodes++;
#if RR_FFMA2
        float c0nx, c0ny, c0fx, c0fy, c1nx, c1ny, c1fx, c1fy, c0nz, c1nz, c0fz, c1fz;
        rr_ffma2(c0nx, c0ny, rr_plane(a.x, nx, magic), rr_plane(a.y, ny, magic), Axy, Bxy);
        rr_ffma2(c0fx, c0fy, rr_plane(a.x, fx, magic), rr_plane(a.y, fy, magic), Axy, Bxy);
        rr_ffma2(c1nx, c1ny, rr_plane(a.w, nx, magic), rr_plane(b.x, ny, magic), Axy, Bxy);
        rr_ffma2(c1fx, c1fy, rr_plane(a.w, fx, magic), rr_plane(b.x, fy, magic), Axy, Bxy);
        rr_ffma2(c0nz, c1nz, rr_plane(a.z, nz, magic), rr_plane(b.y, nz, magic), Azz, Bzz);
        rr_ffma2(c0fz, c1fz, rr_plane(a.z, fz, magic), rr_plane(b.y, fz, magic), Azz, Bzz);
        const float t0n = fmaxf(fmaxf(c0nx, c0ny), fmaxf(c0nz, 0.0f));
        const float t0f = fminf(fminf(c0fx, c0fy), fminf(c0fz, limit));
        const float t1n = fmaxf(fmaxf(c1nx, c1ny), fmaxf(c1nz, 0.0f));
        const float t1f = fminf(fminf(c1fx, c1fy), fminf(c1fz, limit));
#else
        const float t0n = fmaxf(fmaxf(fmaf(rr_plane(a.x, nx, magic), sax, sbx), fmaf(rr_plane(a.y, ny, magic), say, sby)),
                                fmaxf(fmaf(rr_plane(a.z, nz, magic), saz, sbz), 0.0f));
        const float t0f = fminf(fminf(fmaf(rr_plane(a.x, fx, magic), sax, sbx), fmaf(rr_plane(a.y, fy, magic), say, sby)),
                                fminf(fmaf(rr_plane(a.z, fz, magic), saz, sbz), limit));
        const float t1n = fmaxf(fmaxf(fmaf(rr_plane(a.w, nx, magic), sax, sbx), fmaf(rr_plane(b.x, ny, magic), say, sby)),
                                fmaxf(fmaf(rr_plane(b.y, nz, magic), saz, sbz), 0.0f));
        const float t1f = fminf(fminf(fmaf(rr_plane(a.w, fx, magic), sax, sbx), fmaf(rr_plane(b.x, fy, magic), say, sby)),
                                fminf(fmaf(rr_plane(b.y, fz, magic), saz, sbz), limit));
#endif
        const bool h0 = (t0n <= t0f);        /* an absent child (meshes with < 2 leaves) carries an inverted box */
        const bool h1 = (t1n <= t1f);        /* (rr_bvh_pack), which no ray enters: its ref is never followed */
        /* (a select-only form of the four outcomes below — no divergent paths inside the step — measured slower:
         * 1.636 vs 1.594 ms per 16-pose step; the short divergent arms cost less than the extra selects) */
        if (h0 && h1) {
            const bool first0 = t0n <= t1n;
            cur = first0 ? b.z : b.w;
            RR_PUSH(first0 ? b.w : b.z);   /* cannot overflow: the stack holds at most one postponed sibling per level of the
                                              current path and rr_set_mesh rejects a BVH deeper than RR_STACK_SIZE */
        } else if (h0) {
            cur = b.z;
        } else if (h1) {
            cur = b.w;
        } else {
            cur = RR_POP();
        }
    };
#endif
    /* the triangles of the leaf `cur` holds, then the next postponed subtree */
    auto leaf_step = [&]() {
        const uint32_t first = cur & 0x0fffffffu;
        const uint32_t cnt = ((cur >> 28) & 7u) + 1u;
        for (uint32_t k = 0; k < cnt; k++) {
            const float4 q0 = __ldg(tris + 3 * (first + k));
            const float4 q1 = __ldg(tris + 3 * (first + k) + 1);
            const float4 q2 = __ldg(tris + 3 * (first + k) + 2);
            if (STATS) n_tris++;
            float t;
            if (rr_ray_triangle(o, d, rr_v3(q0.x, q0.y, q0.z), rr_v3(q1.x, q1.y, q1.z), rr_v3(q2.x, q2.y, q2.z), tmax, &t)) {
                const int face = (int)__float_as_uint(q0.w);
                if (best_face < 0 || t < best_t || (t == best_t && face < best_face)) {
                    best_t = t; best_face = face; best_slot = (int)(first + k);
                    limit = best_t * 1.00001f + 1e-6f;
                }
            }
        }
        cur = RR_POP();
    };
#if RR_EARLY_EXIT > 0
    for (;;) {
        do {
            if (!(cur & RR_REF_LEAF)) node_step();
        } while (__popc(__ballot_sync(wmask, !(cur & RR_REF_LEAF))) > RR_EARLY_EXIT);
        if ((cur & RR_REF_LEAF) && cur != RR_REF_EMPTY) leaf_step();
        if (!__any_sync(wmask, cur != RR_REF_EMPTY)) break;
    }
#else
    while (cur != RR_REF_EMPTY) {
        while (!(cur & RR_REF_LEAF)) node_step();
        if (cur != RR_REF_EMPTY) leaf_step();
    }
#endif
#undef RR_PUSH
#undef RR_POP
    t_hit = best_t;
    face_hit = best_face;
    return best_slot;
}

/* ------------------------------------------------------------------------------------------------
 * Two rays per lane (RR_DUAL, rr_dual_kernel): the walk of one ray is a chain of dependent steps (node load -> planes ->
 * min/max -> compare -> next ref), and the wide-BVH experiment showed that the chain, not the instruction count, sets
 * its pace. A warp therefore walks TWO groups at once: every lane carries ray A (group 2p) and ray B (group 2p + 1), and
 * one straight-line, branch-free step advances both — two independent chains the scheduler can interleave. A ray that
 * already holds a leaf (or is finished) takes the step on the root node and discards it. Each ray keeps its own order of
 * nodes and triangles, so hits and counters are those of rr_trace.
 * ---------------------------------------------------------------------------------------------- */
struct RRWalk {
    float sax, say, saz, sbx, sby, sbz;
    uint32_t nx, ny, nz;
    float limit;
    uint32_t cur;
    int sp;
    float best_t;
    int best_face, best_slot;
    rr_vec3 o, d;
};

__device__ __forceinline__ void rr_walk_init(RRWalk& W, const float* go, const float* gs, rr_vec3 o, rr_vec3 d, float tmax,
                                             uint32_t root_ref, bool valid, uint32_t* stack)
{
    const float ix = rr_safe_rcp(d.x), iy = rr_safe_rcp(d.y), iz = rr_safe_rcp(d.z);
    W.sax = gs[0] * ix; W.say = gs[1] * iy; W.saz = gs[2] * iz;
    W.sbx = fmaf(-8388608.0f, W.sax, (go[0] - o.x) * ix);
    W.sby = fmaf(-8388608.0f, W.say, (go[1] - o.y) * iy);
    W.sbz = fmaf(-8388608.0f, W.saz, (go[2] - o.z) * iz);
    W.nx = (ix >= 0.f) ? 0x7610u : 0x7632u;
    W.ny = (iy >= 0.f) ? 0x7610u : 0x7632u;
    W.nz = (iz >= 0.f) ? 0x7610u : 0x7632u;
    W.limit = tmax * 1.00001f + 1e-6f;
    W.cur = valid ? root_ref : RR_REF_EMPTY;
    stack[0] = RR_REF_EMPTY;                       /* sentinel: popping it ends the walk */
    W.sp = 1;
    W.best_t = INFINITY; W.best_face = -1; W.best_slot = -1;
    W.o = o; W.d = d;
}

/* one branch-free node step of ray W (a no-op on the ray's state when `inner` is false) */
template <bool STATS>
__device__ __forceinline__ void rr_walk_step(RRWalk& W, const bool inner, const RRNode* __restrict__ nodes, const uint32_t magic,
                                             uint32_t* stack, unsigned& n_nodes)
{
    const uint4* np = reinterpret_cast<const uint4*>(nodes + (inner ? W.cur : 0u));
    const uint4 a = __ldg(np);
    const uint4 b = __ldg(np + 1);
    if (STATS) n_nodes += inner ? 1u : 0u;
    const uint32_t fx = W.nx ^ 0x0022u, fy = W.ny ^ 0x0022u, fz = W.nz ^ 0x0022u;
    const float t0n = fmaxf(fmaxf(fmaf(rr_plane(a.x, W.nx, magic), W.sax, W.sbx), fmaf(rr_plane(a.y, W.ny, magic), W.say, W.sby)),
                            fmaxf(fmaf(rr_plane(a.z, W.nz, magic), W.saz, W.sbz), 0.0f));
    const float t0f = fminf(fminf(fmaf(rr_plane(a.x, fx, magic), W.sax, W.sbx), fmaf(rr_plane(a.y, fy, magic), W.say, W.sby)),
                            fminf(fmaf(rr_plane(a.z, fz, magic), W.saz, W.sbz), W.limit));
    const float t1n = fmaxf(fmaxf(fmaf(rr_plane(a.w, W.nx, magic), W.sax, W.sbx), fmaf(rr_plane(b.x, W.ny, magic), W.say, W.sby)),
                            fmaxf(fmaf(rr_plane(b.y, W.nz, magic), W.saz, W.sbz), 0.0f));
    const float t1f = fminf(fminf(fmaf(rr_plane(a.w, fx, magic), W.sax, W.sbx), fmaf(rr_plane(b.x, fy, magic), W.say, W.sby)),
                            fminf(fmaf(rr_plane(b.y, fz, magic), W.saz, W.sbz), W.limit));
    const bool h0 = (t0n <= t0f), h1 = (t1n <= t1f);
    const bool first0 = (t0n <= t1n);
    const uint32_t near_ref = first0 ? b.z : b.w, far_ref = first0 ? b.w : b.z;      /* when both children are hit */
    const bool both = inner && h0 && h1, none = inner && !(h0 || h1);
    uint32_t nxt = h0 ? (h1 ? near_ref : b.z) : b.w;
    if (both) { stack[W.sp] = far_ref; W.sp++; }
    if (none) { W.sp--; nxt = stack[W.sp]; }
    W.cur = inner ? nxt : W.cur;
}

/* the triangles of the leaf W.cur holds, then the next postponed subtree */
template <bool STATS>
__device__ __forceinline__ void rr_walk_leaf(RRWalk& W, const float4* __restrict__ tris, const float tmax, uint32_t* stack, unsigned& n_tris)
{
    const uint32_t first = W.cur & 0x0fffffffu;
    const uint32_t cnt = ((W.cur >> 28) & 7u) + 1u;
    for (uint32_t k = 0; k < cnt; k++) {
        const float4 q0 = __ldg(tris + 3 * (first + k));
        const float4 q1 = __ldg(tris + 3 * (first + k) + 1);
        const float4 q2 = __ldg(tris + 3 * (first + k) + 2);
        if (STATS) n_tris++;
        float t;
        if (rr_ray_triangle(W.o, W.d, rr_v3(q0.x, q0.y, q0.z), rr_v3(q1.x, q1.y, q1.z), rr_v3(q2.x, q2.y, q2.z), tmax, &t)) {
            const int face = (int)__float_as_uint(q0.w);
            if (W.best_face < 0 || t < W.best_t || (t == W.best_t && face < W.best_face)) {
                W.best_t = t; W.best_face = face; W.best_slot = (int)(first + k);
                W.limit = W.best_t * 1.00001f + 1e-6f;
            }
        }
    }
    W.sp--;
    W.cur = stack[W.sp];
}

template <bool STATS>
__device__ __forceinline__ void rr_trace2(const RRNode* __restrict__ nodes, const float4* __restrict__ tris, uint32_t root_ref,
                                          const float* go, const float* gs, const float tmax,
                                          rr_vec3 oA, rr_vec3 dA, bool validA, rr_vec3 oB, rr_vec3 dB, bool validB,
                                          int2& hitA, int2& hitB, unsigned& n_nodes, unsigned& n_tris)
{
    uint32_t stackA[RR_STACK_SIZE + 1], stackB[RR_STACK_SIZE + 1];
    RRWalk A, B;
    rr_walk_init(A, go, gs, oA, dA, tmax, root_ref, validA, stackA);
    rr_walk_init(B, go, gs, oB, dB, tmax, root_ref, validB, stackB);
    uint32_t magic;
    asm volatile("mov.b32 %0, 0x4B000000;" : "=r"(magic));
    while (A.cur != RR_REF_EMPTY || B.cur != RR_REF_EMPTY) {
        bool inA = !(A.cur & RR_REF_LEAF), inB = !(B.cur & RR_REF_LEAF);
        while (inA || inB) {
            rr_walk_step<STATS>(A, inA, nodes, magic, stackA, n_nodes);
            rr_walk_step<STATS>(B, inB, nodes, magic, stackB, n_nodes);
            inA = !(A.cur & RR_REF_LEAF); inB = !(B.cur & RR_REF_LEAF);
        }
        if (A.cur != RR_REF_EMPTY) rr_walk_leaf<STATS>(A, tris, tmax, stackA, n_tris);
        if (B.cur != RR_REF_EMPTY) rr_walk_leaf<STATS>(B, tris, tmax, stackB, n_tris);
    }
    hitA = make_int2(A.best_slot, __float_as_int(A.best_t));
    hitB = make_int2(B.best_slot, __float_as_int(B.best_t));
}

/* Ken Perlin's reference permutation (image_algorithms.h:14-50 holds it twice back to back) */
__constant__ unsigned char c_perlin_perm[256] = {
    151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,
    247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,68,175,
    74,165,71,134,139,48,27,166,77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,102,143,54,
    65,25,63,161,1,216,80,73,209,76,132,187,208,89,18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,
    52,217,226,250,124,123,5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,
    119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,108,110,79,113,224,232,178,185,112,104,
    218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,49,192,214,31,181,199,106,157,
    184,84,204,176,115,121,50,45,127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180};

/* 2-D slice (z = 0) of the improved-noise function, image_algorithms.h:69-106.
 * With z = 0 the outer blend weight fade(0) is exactly 0, so lerp(w, lower, upper) = lower + 0*(upper-lower)
 * = lower bit-for-bit (upper is finite); only the z-layer-0 gradients are evaluated here.
 * grad(hash, x, y, 0) (image_algorithms.h:108-128) picks u in {x, y}, v in {y, x, 0} by the low 4 hash bits and returns
 * (+-u) + (+-v). Every case is cx * x + cy * y with cx, cy in {-1, -0, +0, +1}: the products are exact and a +-0 term
 * leaves a non-zero partner untouched, so one table lookup + 2 DMUL + 1 DADD replace the chain of predicated selects
 * bit for bit (only the SIGN of an exact zero result can differ, which fabsf() at RadarCPU.cpp:523 removes). */
__device__ __forceinline__ double2 rr_pgrad_coef(int h)
{
    h &= 15;
    const double s1 = (h & 1) ? -1.0 : 1.0, s2 = (h & 2) ? -1.0 : 1.0;
    if (h < 4) return make_double2(s1, s2);                        /* u = x, v = y */
    if (h < 8) return make_double2(s1, s2 * 0.0);                  /* u = x, v = 0 */
    if (h == 12 || h == 14) return make_double2(s2, s1);           /* u = y, v = x */
    return make_double2(s2 * 0.0, s1);                             /* u = y, v = 0 */
}
__device__ __forceinline__ double rr_pfade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
__device__ __forceinline__ double rr_plerp(double t, double a, double b) { return a + t * (b - a); }
/* the part of the evaluation that depends on the second coordinate only: constant along a range column */
struct RRPerlinRow { int Y; double y, ym, v; };
__device__ __forceinline__ RRPerlinRow rr_perlin_row(double sy)
{
    RRPerlinRow r;
    const double fy = floor(sy);
    r.Y = ((int)fy) & 255;
    r.y = sy - fy; r.ym = r.y - 1; r.v = rr_pfade(r.y);
    return r;
}
__device__ __forceinline__ double rr_perlin2(const unsigned char* perm, const double2* grad, double sx, const RRPerlinRow row)
{
    const double fx = floor(sx);
    const int X = ((int)fx) & 255, Y = row.Y;
    const double x = sx - fx, y = row.y;
    const double u = rr_pfade(x), v = row.v;
    const int A = perm[X] + Y, B = perm[(X + 1) & 255] + Y;
    const int AA = perm[A & 255], AB = perm[(A + 1) & 255], BA = perm[B & 255], BB = perm[(B + 1) & 255];
    const double2 c00 = grad[perm[AA] & 15], c10 = grad[perm[BA] & 15], c01 = grad[perm[AB] & 15], c11 = grad[perm[BB] & 15];   /* grad(p[AA], ...): coefficients by the low 4 hash bits */
    const double xm = x - 1, ym = row.ym;
    const double g00 = c00.x * x + c00.y * y, g10 = c10.x * xm + c10.y * y;
    const double g01 = c01.x * x + c01.y * ym, g11 = c11.x * xm + c11.y * ym;
    const double lower = rr_plerp(v, rr_plerp(u, g00, g10), rr_plerp(u, g01, g11));
    return lower + 0.0 * 0.0;   /* == lerp(0, lower, upper) */
}

/* (cell) of a return, RadarCPU.cpp:410-413 */
__device__ __forceinline__ int rr_signal_cell(double time, double resolution)
{
    const float half_time = (float)(time / 2.0);
    const float dist = (float)(0.3 * (double)half_time);
    const double c = (double)dist / resolution;
    return (c >= -2147483648.0 && c < 2147483648.0) ? (int)c : INT32_MIN;
}

/* mono8 conversion of cv::Mat::convertTo(CV_8UC1): cvRound (half to even) + saturate; NaN/overflow -> 0 */
__device__ __forceinline__ uint8_t rr_to_u8(float v)
{
    if (!(fabsf(v) < 2147483648.0f)) return 0;
    const int r = __float2int_rn(v);
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

/* ------------------------------------------------------------------------------------------------
 * Prologue kernels: what is constant per item or per pair of media leaves the per-ray code (rr_internal.h).
 * They call the same rr_detmath.h routines, in the same order, as the per-ray code they replace.
 * ---------------------------------------------------------------------------------------------- */
/* Tam = Tsm * Tas (RadarCPU.cpp:201-206; Tas.t = 0) of every (pose, azimuth) item of a launch sequence, and the map-frame
 * origin of the item's pass-0 rays (wave origin (0,0,0), RadarCPU.cpp:108) */
__global__ void __launch_bounds__(128) rr_prep_kernel(const RRFrameParams P)
{
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (uint32_t)P.n_items) return;
    const int pose_i = (int)(item / (uint32_t)P.az_count);
    const int az = P.az_begin + (int)(item % (uint32_t)P.az_count);
    const rr_pose ps = P.poses[P.pose_per_azimuth ? (pose_i * RR_N_ANGLES + az) : pose_i];
    rr_quat Rsm; Rsm.x = ps.qx; Rsm.y = ps.qy; Rsm.z = ps.qz; Rsm.w = ps.qw;
    const float4 tq = P.tas_quat[az];
    rr_quat Ras; Ras.x = tq.x; Ras.y = tq.y; Ras.z = tq.z; Ras.w = tq.w;
    const rr_quat R = rr_qmul(Rsm, Ras);
    const rr_vec3 T = rr_add(rr_qrot(Rsm, rr_v3(0.f, 0.f, 0.f)), rr_v3(ps.tx, ps.ty, ps.tz));
    const rr_vec3 O0 = rr_add(rr_qrot(R, rr_v3(0.f, 0.f, 0.f)), T);
    P.item_xf[3 * item + 0] = make_float4(R.x, R.y, R.z, R.w);
    P.item_xf[3 * item + 1] = make_float4(T.x, T.y, T.z, 0.f);
    P.item_xf[3 * item + 2] = make_float4(O0.x, O0.y, O0.z, 0.f);
}

/* media-pair constants of Snell/Fresnel (radar_algorithms.h:60-63,80-90,98,110) for every far-side medium of every
 * material table; entry n_mat of a table = "same medium on both sides" (v2 = incidence.velocity, RadarCPU.cpp:277-280) */
__global__ void __launch_bounds__(128) rr_mat_pairs_kernel(const float4* __restrict__ materials, int n_mat, int n_tables, RRMatPair* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tables * (n_mat + 1)) return;
    const int t = i / (n_mat + 1), m = i - t * (n_mat + 1);
    const double wave_v = RR_WAVE_VELOCITY;
    const float v_t = (m < n_mat) ? materials[(size_t)t * n_mat + m].x : (float)wave_v;   /* through a float, RadarCPU.cpp:273 */
    const double n1 = (double)v_t, n2 = wave_v;
    RRMatPair r; r.n1 = n1; r.th_limit = 100.0; r.n12 = 0.0;
    if (n1 > 0.0) {
        const double n21 = n2 / n1;
        if (fabs(n21) <= 1.0) r.th_limit = rr_asin(n21);
        r.n12 = n1 / n2;
    }
    r.rs0 = (n1 - n2) / (n1 + n2);
    out[i] = r;
}

/* ------------------------------------------------------------------------------------------------
 * Kernel 1/3: rr_trace_kernel — ONE PASS of all items: wave -> closest hit -> Snell/Fresnel + BRDF -> returns, children.
 *
 * Wavefront over the whole launch (rr_internal.h): the pass's wave list is cut into groups of 32 consecutive waves; a
 * persistent grid of independent warps pulls groups from a global counter, so every warp round runs 32 live waves
 * whatever the per-azimuth list lengths are (a per-azimuth or per-chunk loop leaves a third of the lanes idle after
 * pass 0). A group appends its surviving children in the reference's order (RadarCPU.cpp:243,290,369: parents in list
 * order, reflection before refraction) by ballot/popc compaction into its own 64 slots; no barrier, no shared memory.
 * ---------------------------------------------------------------------------------------------- */
/* MODE 0: the whole pass in one kernel (rr_trace_kernel). MODE 1 / 2: the pass as two kernels — rr_walk_kernel only casts
 * (closest hit -> hit_rec[j] = (triangle slot, range)) with the small register budget of the walk, rr_shade_kernel reads
 * the hit records and does everything after the cast with the registers the fp64 shading wants. Same code, same order,
 * same bits; which form runs is a host-side switch (rr_api.cu). */
enum { RR_PASS_FUSED = 0, RR_PASS_WALK = 1, RR_PASS_SHADE = 2, RR_PASS_DUAL = 3 };
template <bool STATS, bool DEBUG, int MODE>
__device__ __forceinline__ void rr_pass_body(const RRFrameParams& P, const int pass, uint32_t* s_trace_stack)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t S = (uint32_t)P.n_samples;
    const uint32_t n_in = (pass == 0) ? (uint32_t)P.n_items * S : min(P.pass_total[pass], P.wave_cap);
    const uint32_t n_groups = (n_in + 31u) >> 5;
    const bool last_pass = (pass == P.n_passes - 1);
    const size_t sc = P.slot_cap;
    const int ib = pass & 1, ob = ib ^ 1;
    const float* cf = P.wave_f32 + (size_t)ib * 6 * sc;
    const double* cd = P.wave_f64 + (size_t)ib * 2 * sc;
    const uint32_t* cm = P.wave_mat + (size_t)ib * sc;
    const uint32_t* ci = P.wave_item + (size_t)ib * sc;
    float* nf = P.wave_f32 + (size_t)ob * 6 * sc;
    double* ndp = P.wave_f64 + (size_t)ob * 2 * sc;
    uint32_t* nm = P.wave_mat + (size_t)ob * sc;
    uint32_t* ni = P.wave_item + (size_t)ob * sc;
    const uint32_t* gbase = P.group_base + (size_t)ib * (P.group_cap + 1);
    uint32_t* gcount = P.group_base + (size_t)ob * (P.group_cap + 1);
    uint32_t* item_count_next = P.item_start + (size_t)(pass + 1) * P.item_stride;
    uint32_t* super_next = P.super_count + (size_t)(pass + 1) * P.super_stride;
    uint32_t* item_super_next = P.item_super + (size_t)(pass + 1) * P.item_super_stride;
    int2* sg_cell = P.sig_cell + (size_t)pass * P.wave_cap;
    float2* sg_str = P.sig_strength + (size_t)pass * P.wave_cap;
    const float go[3] = {P.grid_origin[0], P.grid_origin[1], P.grid_origin[2]};
    const float gs[3] = {P.grid_scale[0], P.grid_scale[1], P.grid_scale[2]};
    unsigned stat_nodes = 0, stat_tris = 0;
    uint32_t warp_casts = 0, warp_hits = 0, warp_sigs = 0;

    /* map-frame ray of list position jj (what the cast needs of a wave): RR_PASS_DUAL walks two groups before it shades them */
    auto ray_of = [&](const uint32_t jj, const uint32_t gq, rr_vec3& o_m, rr_vec3& d_m) {
        rr_vec3 wo, wd; uint32_t it;
        if (pass == 0) {
            it = jj / S;
            const uint32_t smp = jj - it * S;
            const float* bd = P.beam_dirs + (P.beam_stride ? (size_t)(it / (uint32_t)P.az_count) * P.beam_stride : (size_t)0);
            wo = rr_v3(0.f, 0.f, 0.f);
            wd = rr_v3(bd[3 * smp], bd[3 * smp + 1], bd[3 * smp + 2]);
        } else {
            uint32_t gg = P.first_src[gq];
            while (gbase[gg + 1] <= jj) gg++;
            const size_t slot = (size_t)gg * 64u + (jj - gbase[gg]);
            wo = rr_v3(__ldg(cf + 0 * sc + slot), __ldg(cf + 1 * sc + slot), __ldg(cf + 2 * sc + slot));
            wd = rr_v3(__ldg(cf + 3 * sc + slot), __ldg(cf + 4 * sc + slot), __ldg(cf + 5 * sc + slot));
            it = __ldg(ci + slot);
        }
        const float4 xr = __ldg(P.item_xf + 3 * (size_t)it), xt = __ldg(P.item_xf + 3 * (size_t)it + 1);
        rr_quat R; R.x = xr.x; R.y = xr.y; R.z = xr.z; R.w = xr.w;
        if (pass == 0) { const float4 x0 = __ldg(P.item_xf + 3 * (size_t)it + 2); o_m = rr_v3(x0.x, x0.y, x0.z); }
        else o_m = rr_add(RR_QROT(R, wo), rr_v3(xt.x, xt.y, xt.z));
        d_m = RR_QROT(R, wd);
    };
    uint32_t pair_g = 0, pair_left = 0;
    while (true) {
        uint32_t g = 0;
        if (MODE == RR_PASS_DUAL) {
            if (pair_left == 0u) {                         /* next pair of groups: walk both, then shade them one by one */
                if (lane == 0) pair_g = atomicAdd(P.work_counter + pass, 2u);
                pair_g = __shfl_sync(RR_FULL, pair_g, 0);
                if (pair_g >= n_groups) break;
                const uint32_t jA = pair_g * 32u + (uint32_t)lane, jB = jA + 32u;
                const bool vA = jA < n_in, vB = jB < n_in;
                rr_vec3 oA = rr_v3(0.f, 0.f, 0.f), dA = rr_v3(1.f, 0.f, 0.f), oB = oA, dB = dA;
                if (vA) ray_of(jA, pair_g, oA, dA);
                if (vB) ray_of(jB, pair_g + 1u, oB, dB);
                int2 hA, hB;
                rr_trace2<STATS>(P.nodes, P.tris, P.root_ref, go, gs, 1000.0f, oA, dA, vA, oB, dB, vB, hA, hB, stat_nodes, stat_tris);
                if (vA) P.hit_rec[jA] = hA;                /* read back by this very thread below */
                if (vB) P.hit_rec[jB] = hB;
                pair_left = 2u;
            }
            g = pair_g + (2u - pair_left);
            pair_left--;
            if (g >= n_groups) continue;
        } else {
            if (lane == 0) g = atomicAdd(P.work_counter + (MODE == RR_PASS_SHADE ? RR_MAX_PASSES + 1 : 0) + pass, 1u);
            g = __shfl_sync(RR_FULL, g, 0);
            if (g >= n_groups) break;
        }
        const uint32_t j = g * 32u + (uint32_t)lane;
        const bool active = j < n_in;
        const uint32_t act_mask = __ballot_sync(RR_FULL, active);

        uint32_t item = 0;
        bool keep0 = false, keep1 = false, hit = false;
        uint32_t n_child = 0, n_sig = 0;
        rr_vec3 c_o = rr_v3(0, 0, 0), c_d0 = c_o, c_d1 = c_o;
        double c_time = 0, c_e0 = 0, c_e1 = 0;
        uint32_t c_m0 = 0, c_m1 = 0;
        int sig_cell0 = INT32_MIN, sig_cell1 = INT32_MIN;
        float sig_s0 = 0, sig_s1 = 0, sig_t0 = 0, sig_t1 = 0;
        int face = -1, az = 0; float range = 0.f; float dbg_energy = 0.f;

        if (active) {
            RRWave w;
            if (pass == 0) {                              /* RadarCPU.cpp:106-114,184 */
                item = j / S;
                const uint32_t smp = j - item * S;
                w.o = rr_v3(0.f, 0.f, 0.f);
                const float* bd = P.beam_dirs + (P.beam_stride ? (size_t)(item / (uint32_t)P.az_count) * P.beam_stride : (size_t)0);   /* per-goal bundles: rr_gen_radar_images */
                w.d = rr_v3(bd[3 * smp], bd[3 * smp + 1], bd[3 * smp + 2]);
                w.energy = 1.0; w.time = 0.0; w.mat = 0u;
            } else {                                      /* list position j -> slot (rr_internal.h) */
                uint32_t gg = P.first_src[g];
                while (gbase[gg + 1] <= j) gg++;
                const size_t slot = (size_t)gg * 64u + (j - gbase[gg]);
                /* wave state is read once and written once per pass: streaming loads/stores (evict-first) keep the BVH
                 * nodes and the traversal stacks in L1/L2 instead */
                w.o = rr_v3(__ldcs(cf + 0 * sc + slot), __ldcs(cf + 1 * sc + slot), __ldcs(cf + 2 * sc + slot));
                w.d = rr_v3(__ldcs(cf + 3 * sc + slot), __ldcs(cf + 4 * sc + slot), __ldcs(cf + 5 * sc + slot));
                w.energy = __ldcs(cd + 0 * sc + slot); w.time = __ldcs(cd + 1 * sc + slot);
                w.mat = __ldcs(cm + slot);
                item = __ldcs(ci + slot);
            }
            dbg_energy = (float)w.energy;
            /* Tam = Tsm * Tas of the item (rr_prep_kernel) */
            const float4 xr = __ldg(P.item_xf + 3 * (size_t)item), xt = __ldg(P.item_xf + 3 * (size_t)item + 1);
            rr_quat R; R.x = xr.x; R.y = xr.y; R.z = xr.z; R.w = xr.w;
            const uint32_t pose_i = (P.material_stride | P.mat_pair_stride) ? item / (uint32_t)P.az_count : 0u;
            if (DEBUG) az = P.az_begin + (int)(item % (uint32_t)P.az_count);
            /* ray into the map frame; closest hit within [0, 1000] m (radar_algorithms.cpp:157-158) */
            rr_vec3 o_m;
            if (pass == 0) { const float4 x0 = __ldg(P.item_xf + 3 * (size_t)item + 2); o_m = rr_v3(x0.x, x0.y, x0.z); }
            else o_m = rr_add(RR_QROT(R, w.o), rr_v3(xt.x, xt.y, xt.z));
            const rr_vec3 d_m = RR_QROT(R, w.d);
            int slot_t;
            if (MODE == RR_PASS_SHADE || MODE == RR_PASS_DUAL) {
                const int2 hr = (MODE == RR_PASS_DUAL) ? P.hit_rec[j] : __ldcs(P.hit_rec + j);
                slot_t = hr.x; range = __int_as_float(hr.y);
                if (DEBUG && slot_t >= 0) face = (int)__float_as_uint(__ldg(P.tris + 3 * slot_t).w);
            } else {
                slot_t = rr_trace<STATS>(P.nodes, P.tris, P.root_ref, go, gs, o_m, d_m, 1000.0f,
                                         range, face, stat_nodes, stat_tris, s_trace_stack, act_mask);
            }
            if (MODE == RR_PASS_WALK) __stcs(P.hit_rec + j, make_int2(slot_t, __float_as_int(range)));
            if (MODE != RR_PASS_WALK && slot_t >= 0) {
                const float4 q1 = __ldg(P.tris + 3 * slot_t + 1);
                const float4 q2 = __ldg(P.tris + 3 * slot_t + 2);
                const uint32_t obj = __float_as_uint(q1.w);
                if (obj >= (uint32_t)P.n_objects) {
                    atomicExch(&P.error_flags[1], 1);
                } else {
                    hit = true;
                    /* geometric normal -> ray frame, facing the ray, re-normalised (RadarCPU.cpp:248) */
                    rr_vec3 n = RR_NORMALIZE(rr_cross(rr_v3(q1.x, q1.y, q1.z), rr_v3(q2.x, q2.y, q2.z)));
                    n = RR_QROT(rr_qinv(R), n);
                    if (rr_dot(w.d, n) > 0.0f) n = rr_neg(n);
                    n = RR_NORMALIZE(n);

                    /* move to the surface (radar_types.h:108-113) */
                    const rr_vec3 p_hit = rr_add(w.o, rr_muls(w.d, range));
                    const double wave_v = RR_WAVE_VELOCITY;
                    const double t_hit = w.time + RR_DDIV((double)range, wave_v);

                    /* medium on the far side (RadarCPU.cpp:266-280) and its Snell/Fresnel constants (rr_mat_pairs_kernel) */
                    const uint32_t air = (uint32_t)P.material_id_air;
                    const uint32_t mat_t = (w.mat == air) ? (uint32_t)P.object_materials[obj] : air;
                    const double2* mpp = reinterpret_cast<const double2*>(P.mat_pairs + (size_t)pose_i * P.mat_pair_stride
                                                                           + ((w.mat != mat_t) ? mat_t : (uint32_t)P.n_materials));
                    const double2 mp0 = __ldg(mpp), mp1 = __ldg(mpp + 1);          /* (n1, th_limit) (n12, rs0) */

                    /* Snell/Fresnel (radar_algorithms.h:55-139): n1 := v2, n2 := v1 */
                    const double n1 = mp0.x;
                    const float cos_i = rr_dot(rr_neg(w.d), n);
                    const float th_if = rr_acosf(cos_i);
                    const double th_i = (double)th_if;
                    const rr_vec3 d_refl = rr_add(w.d, rr_muls(rr_muls(n, 2.0f), rr_dot(rr_neg(n), w.d)));
                    rr_vec3 d_refr = rr_v3(0.f, 0.f, 0.f);
                    /* refraction angle acos(refr . (-n)) (:106). Without a refracted ray (total reflection, v2 = 0) the
                     * dot product of the zero vector is +-0 whatever n is, so the angle is the constant (float)acos(0)
                     * (a NaN in n has already made th_i NaN and with it everything below). */
                    double th_t = (double)RR_ACOSF_OF_ZERO;
                    if (n1 > 0.0 && th_i <= mp0.y) {
                        rr_vec3 nn = n;
                        if (rr_dot(nn, w.d) > 0.0f) nn = rr_neg(nn);
                        const double n12 = mp1.x;                                  /* n2 = 0.3 > 0 always */
                        const double c = rr_cos(th_i);
                        const double k = n12 * c - sqrt(1 - n12 * n12 * (1 - c * c));
                        d_refr = rr_add(rr_muls(w.d, (float)n12), rr_muls(nn, (float)k));
                        th_t = (double)rr_acosf(rr_dot(d_refr, rr_neg(nn)));
                    }
                    double rs, rp;
                    const double th_sum = th_i + th_t;
                    if (th_sum < 0.0001) {
                        rs = mp1.y; rp = rs;
                    } else if (th_sum > M_PI - 0.0001) {
                        rs = 1.0; rp = 1.0;
                    } else {
                        const rr_sincos_t pd = rr_sincos_parts(th_i - th_t), psum = rr_sincos_parts(th_sum);
                        rs = RR_DDIV(-rr_sin_of(pd), rr_sin_of(psum));
                        rp = RR_DDIV(rr_tan_of_ol(pd), rr_tan_of_ol(psum));
                    }
                    const double Reff = 0.5 * (rs * rs) + (1.0 - 0.5) * (rp * rp);
                    const double Teff = 1.0 - Reff;
                    const double e_refl = Reff * w.energy;
                    const double e_refr = Teff * w.energy;
                    const double thr = (double)0.001f;                 /* Radar.cpp:24 */

                    c_o = p_hit; c_time = t_hit;
                    if (e_refl > thr) {                                /* RadarCPU.cpp:288 */
                        keep0 = true; c_d0 = d_refl; c_e0 = e_refl; c_m0 = w.mat;
                        if (w.mat == air) {                            /* :302 — return to the sensor */
                            const float e_f = (float)e_refl;
                            const float4 mt = __ldg(P.materials + (size_t)pose_i * P.material_stride + mat_t);
                            if (pass == 0 || P.record_multi_reflection) {
                                /* BRDF, radar_algorithms.h:168-187: (A, B, C) = (ambient, diffuse, specular) */
                                const float ret = rr_brdf(th_if, mt) * e_f;
                                const float t_back = (float)(t_hit * 2.0);
                                sig_cell0 = rr_signal_cell((double)t_back, P.resolution);
                                sig_s0 = ret; sig_t0 = t_back; n_sig = 1;
                            }
                            if (pass > 0 && P.record_multi_path) {     /* :325-360 */
                                const float dist_f = rr_l2norm(p_hit);
                                const rr_vec3 to_hit = rr_divs(p_hit, rr_l2norm(p_hit));
                                const double t_sensor = RR_DDIV((double)dist_f, wave_v);
                                const double view = (double)rr_dot(w.d, to_hit);
                                const float ang = rr_acosf(rr_dot(rr_neg(d_refl), to_hit));
                                if (view > P.multipath_threshold) {
                                    const float ret = rr_brdf(ang, mt) * e_f;
                                    const double t_air = t_hit + t_sensor;
                                    const int cell = rr_signal_cell(t_air, P.resolution);
                                    if (n_sig == 0) { sig_cell0 = cell; sig_s0 = ret; sig_t0 = (float)t_air; }
                                    else { sig_cell1 = cell; sig_s1 = ret; sig_t1 = (float)t_air; }
                                    n_sig++;
                                }
                            }
                        }
                    }
                    if (e_refr > thr) {                                /* :364-370 */
                        keep1 = true; c_d1 = d_refr; c_e1 = e_refr; c_m1 = mat_t;
                    }
                    n_child = (keep0 ? 1u : 0u) + (keep1 ? 1u : 0u);
                }
            }
            /* returns of wave j, in the order RadarCPU.cpp:322,358 appends them; a slot without a return (and a return
             * whose time is not a number, which can not reach a bin either) carries cell INT32_MIN */
            if (MODE != RR_PASS_WALK) {
                __stcs(sg_cell + j, make_int2((n_sig >= 1) ? sig_cell0 : INT32_MIN, (n_sig >= 2) ? sig_cell1 : INT32_MIN));
                if (n_sig) __stcs(sg_str + j, make_float2(sig_s0, sig_s1));
            }
            if (DEBUG && MODE != RR_PASS_WALK) {
                const size_t r0 = (size_t)pass * P.wave_cap + j;
                rr_cast_record r; r.azimuth = az; r.pass = pass; r.face_id = hit ? face : -1;
                r.range = hit ? range : 0.f; r.energy = dbg_energy; r.n_children = (int)n_child;
                P.dbg_casts[r0] = r;
                rr_signal_record s0; s0.azimuth = (n_sig >= 1) ? az : -1; s0.cell = sig_cell0; s0.strength = sig_s0; s0.time = sig_t0;
                rr_signal_record s1; s1.azimuth = (n_sig >= 2) ? az : -1; s1.cell = sig_cell1; s1.strength = sig_s1; s1.time = sig_t1;
                P.dbg_signals[2 * r0] = s0; P.dbg_signals[2 * r0 + 1] = s1;
            }
        }

        if (MODE == RR_PASS_WALK) continue;            /* the walk kernel only casts */
        /* ordered compaction inside the group (reflection before refraction, parents in list order).
         * The reference also builds waves_new in its last pass and drops it (RadarCPU.cpp:380-389):
         * nothing traces those waves, so the last pass appends none (n_children is still reported). */
        warp_casts += __popc(act_mask);
        warp_hits += __popc(__ballot_sync(RR_FULL, hit));
        warp_sigs += __popc(__ballot_sync(RR_FULL, n_sig >= 1)) + __popc(__ballot_sync(RR_FULL, n_sig >= 2));
        if (!last_pass) {
            if (P.pose_passes && active && pass + 1 >= P.pose_passes[item / (uint32_t)P.az_count]) { keep0 = false; keep1 = false; }   /* this goal's last pass */
            const uint32_t m0 = __ballot_sync(RR_FULL, keep0), m1 = __ballot_sync(RR_FULL, keep1);
            size_t co = (size_t)g * 64u + __popc(m0 & lt_mask) + __popc(m1 & lt_mask);
            const float skip = 0.001f;                                     /* RadarCPU.cpp:374-378 */
            if (keep0) {
                const rr_vec3 o2 = rr_add(c_o, rr_muls(c_d0, skip));
                __stcs(nf + 0 * sc + co, o2.x); __stcs(nf + 1 * sc + co, o2.y); __stcs(nf + 2 * sc + co, o2.z);
                __stcs(nf + 3 * sc + co, c_d0.x); __stcs(nf + 4 * sc + co, c_d0.y); __stcs(nf + 5 * sc + co, c_d0.z);
                __stcs(ndp + 0 * sc + co, c_e0); __stcs(ndp + 1 * sc + co, c_time + (double)skip / RR_WAVE_VELOCITY);
                __stcs(nm + co, c_m0); __stcs(ni + co, item);
                co++;
            }
            if (keep1) {
                const rr_vec3 o2 = rr_add(c_o, rr_muls(c_d1, skip));
                __stcs(nf + 0 * sc + co, o2.x); __stcs(nf + 1 * sc + co, o2.y); __stcs(nf + 2 * sc + co, o2.z);
                __stcs(nf + 3 * sc + co, c_d1.x); __stcs(nf + 4 * sc + co, c_d1.y); __stcs(nf + 5 * sc + co, c_d1.z);
                __stcs(ndp + 0 * sc + co, c_e1); __stcs(ndp + 1 * sc + co, c_time + (double)skip / RR_WAVE_VELOCITY);
                __stcs(nm + co, c_m1); __stcs(ni + co, item);
            }
            if (lane == 0) {                           /* group count + its 1024-group partial sum (rr_scan_kernel) */
                const uint32_t c = __popc(m0) + __popc(m1);
                gcount[g] = c;
                if (c) atomicAdd(super_next + (g >> 10), c);
            }
            if (active) {                              /* the next list's per-item lengths (items are contiguous runs) */
                const uint32_t peers = __match_any_sync(act_mask, item);
                const uint32_t c = __popc(m0 & peers) + __popc(m1 & peers);
                if ((peers & lt_mask) == 0u && c) { atomicAdd(item_count_next + item, c); atomicAdd(item_super_next + (item >> 10), c); }
            }
        }
    }
    if (lane == 0 && warp_casts) {
        atomicAdd(&P.counters[0], (unsigned long long)warp_casts);
        atomicAdd(&P.counters[1], (unsigned long long)warp_hits);
        atomicAdd(&P.counters[2], (unsigned long long)warp_sigs);
    }
    if (STATS) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            stat_nodes += __shfl_xor_sync(RR_FULL, stat_nodes, off);
            stat_tris += __shfl_xor_sync(RR_FULL, stat_tris, off);
        }
        if (lane == 0) {
            atomicAdd(&P.counters[3], (unsigned long long)stat_nodes);
            atomicAdd(&P.counters[4], (unsigned long long)stat_tris);
        }
    }
}

template <bool STATS, bool DEBUG>
__global__ void __launch_bounds__(RR_TRACE_BLOCK, RR_MIN_BLOCKS) rr_trace_kernel(const RRFrameParams P, const int pass)
{
    extern __shared__ uint32_t s_trace_stack[];
    rr_pass_body<STATS, DEBUG, RR_PASS_FUSED>(P, pass, s_trace_stack);
}
#ifndef RR_WALK_MIN_BLOCKS
#define RR_WALK_MIN_BLOCKS 12      /* resident 128-thread CTAs per SM of the cast-only kernel (40 registers) */
#endif
#ifndef RR_SHADE_MIN_BLOCKS
#define RR_SHADE_MIN_BLOCKS 8      /* ... of the shading kernel (64 registers) */
#endif
template <bool STATS>
__global__ void __launch_bounds__(RR_TRACE_BLOCK, RR_WALK_MIN_BLOCKS) rr_walk_kernel(const RRFrameParams P, const int pass)
{
    extern __shared__ uint32_t s_trace_stack[];
    rr_pass_body<STATS, false, RR_PASS_WALK>(P, pass, s_trace_stack);
}
#ifndef RR_DUAL_MIN_BLOCKS
#define RR_DUAL_MIN_BLOCKS 6       /* resident 128-thread CTAs per SM of the two-rays-per-lane kernel (80 registers) */
#endif
template <bool STATS, bool DEBUG>
__global__ void __launch_bounds__(RR_TRACE_BLOCK, RR_DUAL_MIN_BLOCKS) rr_dual_kernel(const RRFrameParams P, const int pass)
{
    rr_pass_body<STATS, DEBUG, RR_PASS_DUAL>(P, pass, nullptr);
}
template <bool DEBUG>
__global__ void __launch_bounds__(RR_TRACE_BLOCK, RR_SHADE_MIN_BLOCKS) rr_shade_kernel(const RRFrameParams P, const int pass)
{
    rr_pass_body<false, DEBUG, RR_PASS_SHADE>(P, pass, nullptr);
}

/* ------------------------------------------------------------------------------------------------
 * Kernel 2/3: rr_scan_kernel — between two passes: child counts of the groups of pass `pass - 1` -> exclusive prefix
 * group_base[] (+ total = list length of `pass`), first_src[] of the new list's groups, and the per-item counts ->
 * item_start[pass][]. One CTA per 1024 counts; the offset of a CTA is the sum of the 1024-count partial sums the trace
 * kernel accumulated (super_count / item_super), so there is no chained wait between CTAs. Order-preserving by
 * construction (prefix sums), so the new list is the reference's.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t rr_block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const uint32_t nb = __shfl_up_sync(RR_FULL, incl, off); if (lane >= off) incl += nb; }
    __syncthreads();                                   /* s_warp may still be read from the previous call */
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t wv = (lane < RR_SCAN_BLOCK / 32) ? s_warp[lane] : 0u;
        uint32_t wi = wv;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const uint32_t nb = __shfl_up_sync(RR_FULL, wi, off); if (lane >= off) wi += nb; }
        if (lane < RR_SCAN_BLOCK / 32) s_warp[lane] = wi - wv;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[wid] + incl - v;
}

__global__ void __launch_bounds__(RR_SCAN_BLOCK) rr_scan_kernel(const RRFrameParams P, const int pass)
{
    __shared__ uint32_t s_warp[33];
    const uint32_t tid = threadIdx.x, b = blockIdx.x;
    const uint32_t S = (uint32_t)P.n_samples;
    const uint32_t n_prev = (pass == 1) ? (uint32_t)P.n_items * S : min(P.pass_total[pass - 1], P.wave_cap);
    const uint32_t n_groups = (n_prev + 31u) >> 5;
    const uint32_t n_super = (n_groups + RR_SCAN_BLOCK - 1) / RR_SCAN_BLOCK;
    uint32_t* gb = P.group_base + (size_t)(pass & 1) * (P.group_cap + 1);
    /* ---- groups: CTA b owns groups [1024 b, 1024 b + 1024); the trace kernel already summed them into super_count[b] */
    if (b < n_super) {
        const uint32_t* sc = P.super_count + (size_t)pass * P.super_stride;
        uint32_t part = 0, before, dummy;
        for (uint32_t i = tid; i < b; i += RR_SCAN_BLOCK) part += sc[i];
        rr_block_exclusive_scan(part, s_warp, &before);            /* before = children of all earlier CTAs' groups */
        const uint32_t g = b * RR_SCAN_BLOCK + tid;
        const uint32_t c = (g < n_groups) ? gb[g] : 0u;
        uint32_t tile_total;
        const uint32_t base = before + rr_block_exclusive_scan(c, s_warp, &tile_total);
        if (g < n_groups) {
            gb[g] = base;
            for (uint32_t go = (base + 31u) >> 5; go * 32u < base + c && go <= P.group_cap; go++) P.first_src[go] = g;
        }
        if (b == n_super - 1 && tid == 0) {
            const uint32_t total = before + tile_total;
            gb[n_groups] = total;                       /* sentinel: ends the slot walk of the trace kernel */
            P.pass_total[pass] = total;
            if (total > P.wave_cap) atomicExch(&P.error_flags[0], 1);
        }
        (void)dummy;
    } else if (b == 0 && tid == 0) {                    /* empty previous list */
        gb[0] = 0u; P.pass_total[pass] = 0u;
    }
    /* ---- items: CTA b owns items [1024 b, 1024 b + 1024) */
    const uint32_t n_items = (uint32_t)P.n_items;
    const uint32_t n_isuper = (n_items + RR_SCAN_BLOCK - 1) / RR_SCAN_BLOCK;
    if (b < n_isuper) {
        uint32_t* is = P.item_start + (size_t)pass * P.item_stride;
        const uint32_t* isc = P.item_super + (size_t)pass * P.item_super_stride;
        uint32_t part = 0, before;
        for (uint32_t i = tid; i < b; i += RR_SCAN_BLOCK) part += isc[i];
        rr_block_exclusive_scan(part, s_warp, &before);
        const uint32_t it = b * RR_SCAN_BLOCK + tid;
        const uint32_t c = (it < n_items) ? is[it] : 0u;
        uint32_t tile_total;
        const uint32_t base = before + rr_block_exclusive_scan(c, s_warp, &tile_total);
        if (it < n_items) is[it] = base;
        if (b == n_isuper - 1 && tid == 0) is[n_items] = before + tile_total;
        uint32_t longest = c;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) longest = max(longest, __shfl_xor_sync(RR_FULL, longest, off));
        if ((tid & 31u) == 0u) atomicMax(&P.counters[5], (unsigned long long)max(longest, (pass == 1) ? S : 0u));
    }
}

/* ------------------------------------------------------------------------------------------------
 * Kernel 3/3: rr_draw_kernel — returns -> range column (RadarCPU.cpp:402-450) -> energy_max, ambient noise,
 * normalise, mono8 (RadarCPU.cpp:453-542). One CTA per (pose, azimuth); the column lives in shared memory.
 *
 * The reference adds every return's window of W weighted bins in list order, so a bin's float value depends on the ORDER
 * of the additions that reach it (and on nothing else: bins are independent). The kernel keeps that order per bin without
 * atomics and without replaying the list:
 *   1. count   — the return list is cut into 64 contiguous pieces; every return adds 1 to the counter (piece, granule)
 *                of each 32-bin granule its window overlaps (packed 16-bit counters, shared-memory atomics, any order);
 *   2. scan    — per granule: prefix over the pieces; over the granules: offsets of the per-granule entry lists;
 *   3. fill    — thread p (of the first 64) walks piece p front to back and appends (window start, strength) to the lists of the
 *                granules it overlaps through its OWN cursors: list order = order inside every granule list, no
 *                synchronisation, no ranking;
 *   4. add     — a warp takes a granule (dynamic queue), lane <-> bin, the bin's value sits in a REGISTER while the lane
 *                walks the granule's list front to back: one broadcast shared-memory load per entry and the dependent
 *                add chain of the reference, branch-free. A window only engages the lanes of the granules it overlaps;
 *                no return is ever looked at by a warp that has no bin in its window.
 * Lists longer than the entry buffer (many passes / wide kernels) are processed in chunks of the return list; the column
 * stays in shared memory between chunks.
 * Epilogue per cell, then the mono8 column goes out as
 *   RR_OUT_GROUP    row-major image (the reference's cv::Mat, Radar.cpp:34) in 8-byte row segments: the column is staged in
 *                   global memory (16-byte stores, L2-resident) and the CTA that finishes LAST among 8 adjacent azimuths
 *                   transposes the 8 columns — nobody waits for anybody;
 *   RR_OUT_CLUSTER  the same through a thread-block cluster of 8 CTAs reading each other's columns over distributed
 *                   shared memory (RR_DRAW_USE_CLUSTER=1; measured slower: the 8 CTAs wait for their slowest column);
 *   RR_OUT_BYTES    byte stores at stride 400 (shards whose column range is not a multiple of 8, odd scroll, debug);
 *   RR_OUT_COLUMNS  column-major shard / NVLink peer stores (16-byte vectors).
 * ---------------------------------------------------------------------------------------------- */
#ifndef RR_DRAW_RETURNS
#define RR_DRAW_RETURNS 1024          /* returns of a chunk staged in shared memory: 8 KB */
#endif
#ifndef RR_DRAW_ENTRIES
#define RR_DRAW_ENTRIES 3072          /* 16-bit list entries of a chunk: 6 KB (W = 35: <= 3 granules per return) */
#endif
#ifndef RR_DRAW_PIECES
#define RR_DRAW_PIECES 64             /* pieces of the return list = threads that fill the lists (<= RR_BLOCK); measured 32 -> 0.537 ms, 64 -> 0.523 */
#endif
#define RR_DRAW_GROUP 8               /* adjacent azimuths per row segment */
enum { RR_OUT_GROUP = 0, RR_OUT_BYTES = 1, RR_OUT_COLUMNS = 2, RR_OUT_CLUSTER = 3 };

/* 8 byte columns (4 cells each) -> 4 row segments of 8 bytes */
__device__ __forceinline__ void rr_store_segments(const uint32_t x[RR_DRAW_GROUP], uint8_t* seg, int cell0, int C)
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (cell0 + i < C) {
            const uint32_t sel = (uint32_t)i | ((uint32_t)(4 + i) << 4);
            const uint32_t lo = __byte_perm(__byte_perm(x[0], x[1], sel), __byte_perm(x[2], x[3], sel), 0x5410);
            const uint32_t hi = __byte_perm(__byte_perm(x[4], x[5], sel), __byte_perm(x[6], x[7], sel), 0x5410);
            *reinterpret_cast<uint2*>(seg + (size_t)(cell0 + i) * RR_N_ANGLES) = make_uint2(lo, hi);
        }
    }
}

#ifndef RR_DRAW_MIN_CTAS
#define RR_DRAW_MIN_CTAS 5       /* 256-thread CTAs per SM (measured: 4 -> 0.591 ms, 5 -> 0.561, 6 -> 0.572) */
#endif
template <bool DEBUG, int OUT>
__global__ void __launch_bounds__(RR_BLOCK, RR_DRAW_MIN_CTAS) rr_draw_kernel(const RRFrameParams P)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ double s_weights[RR_MAX_DENOISE];         /* float weights widened once (the splat multiplies in double) */
    __shared__ unsigned char s_perm[256];
    __shared__ double2 s_grad[16];                       /* gradient coefficients by the low 4 hash bits */
    __shared__ float s_red[RR_WARPS];
    __shared__ uint32_t s_begin[RR_MAX_PASSES], s_voff[RR_MAX_PASSES + 1];
    __shared__ uint32_t s_next, s_nne, s_last;
    __shared__ uint32_t s_plain;                         /* 1 while every weight and every strength of the chunk is finite and >= 0 */

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t item = blockIdx.x;
    const int pose_i = (int)(item / (uint32_t)P.az_count);
    const int az = P.az_begin + (int)(item % (uint32_t)P.az_count);
    const int C = P.n_cells, C16 = (C + 15) & ~15;
    const int n_passes = P.n_passes;
    const int n_gran = (C + 31) >> 5, G1 = (n_gran + 2) & ~1;    /* <= 313 granules for n_cells <= 10000; G1 even */
    float* s_col = reinterpret_cast<float*>(s_raw);                                    /* [C16] this azimuth's range column     */
    uint2* s_ret = reinterpret_cast<uint2*>(s_raw + (size_t)C16 * 4);                  /* [RR_DRAW_RETURNS] (cell, strength)    */
    uint16_t* s_ent = reinterpret_cast<uint16_t*>(s_ret + RR_DRAW_RETURNS);            /* [RR_DRAW_ENTRIES] return indices      */
    uint8_t* s_bytes = reinterpret_cast<uint8_t*>(s_ret);                              /* [C16] mono8 column, after the lists   */
    uint16_t* s_tab = s_ent + RR_DRAW_ENTRIES;                                         /* [PIECES][G1] counts -> piece prefixes */
    uint32_t* s_off = reinterpret_cast<uint32_t*>(s_tab + RR_DRAW_PIECES * G1);        /* [G1] list offsets                     */
    uint16_t* s_ne = reinterpret_cast<uint16_t*>(s_off + G1);                          /* [G1] granules with a non-empty list   */

    bool w_plain = true;
    for (int i = tid; i < RR_MAX_DENOISE; i += RR_BLOCK) {
        const float wv = (i < P.denoise_width) ? P.denoise_weights[i] : 0.0f;
        s_weights[i] = (double)wv;
        w_plain = w_plain && (wv >= 0.0f) && (wv < INFINITY);
    }
    const bool weights_plain = __syncthreads_and(w_plain) != 0;
    if (tid < 64) reinterpret_cast<uint32_t*>(s_perm)[tid] = reinterpret_cast<const uint32_t*>(c_perlin_perm)[tid];
    if (tid < 16) s_grad[tid] = rr_pgrad_coef(tid);
    for (int i = tid; i < (C16 >> 2); i += RR_BLOCK) reinterpret_cast<float4*>(s_col)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {                                                /* this item's run of every pass list, concatenated */
        const uint32_t S = (uint32_t)P.n_samples;
        uint32_t run = 0;
        for (int p = 0; p < n_passes; p++) {
            uint32_t b, e, lim;
            if (p == 0) { b = item * S; e = b + S; lim = (uint32_t)P.n_items * S; }
            else {
                const uint32_t* is = P.item_start + (size_t)p * P.item_stride;
                b = is[item]; e = is[item + 1]; lim = min(P.pass_total[p], P.wave_cap);
            }
            b = min(b, lim); e = max(b, min(e, lim));
            s_begin[p] = b; s_voff[p] = run; run += e - b;
        }
        s_voff[n_passes] = run;
    }
    __syncthreads();
    const uint32_t n_slots = s_voff[n_passes];

    const int W = P.denoise_on ? P.denoise_width : 1;
    const int mode = P.denoise_on ? P.denoise_mode : 0;
    const int lo_bin = P.denoise_on ? 1 : 0;                       /* glob_id > 0 (:424) only with denoising */
    const uint32_t max_gr = (uint32_t)(W + 30) / 32u + 1u;         /* granules one window can overlap */
    /* a wave carries up to two returns (path, multipath: RadarCPU.cpp:322,358); without record_multi_path the second slot
     * is never filled and is not even looked at */
    const uint32_t rpw = P.record_multi_path ? 2u : 1u;
    /* waves per chunk: their returns fit the staging table and, whatever the windows, their entries fit the lists */
    const uint32_t ch = max(1u, min((uint32_t)RR_DRAW_RETURNS, (uint32_t)RR_DRAW_ENTRIES / max_gr) / rpw);

    /* window of a return: bins [max(st, lo_bin), min(st + W, C)), st = cell - mode; cell < C (:414); very negative cells
     * (no return, time = -inf/NaN) can not reach a bin */
    auto window = [&](int cell, int& g_lo, int& g_hi) -> bool {
        if (!(cell < C) || !(cell > -RR_MAX_DENOISE - 1)) return false;
        const int st = cell - mode;
        const int lo = max(st, lo_bin), hi = min(st + W, C);
        if (lo >= hi) return false;
        g_lo = lo >> 5; g_hi = (hi - 1) >> 5;
        return true;
    };

    float m = 0.0f;                                                /* running max_val (:428-431), per lane */
    for (uint32_t c0 = 0; c0 < n_slots; c0 += ch) {
        const uint32_t n_w = min(ch, n_slots - c0), n_ret = rpw * n_w;
        const uint32_t plen = (n_ret + RR_DRAW_PIECES - 1) / RR_DRAW_PIECES;      /* returns per piece */
        /* ---- 0. clear the counters */
        for (int i = tid; i < (RR_DRAW_PIECES * G1) / 2; i += RR_BLOCK) reinterpret_cast<uint32_t*>(s_tab)[i] = 0u;
        if (tid == 0) { s_next = 0u; s_plain = weights_plain ? 1u : 0u; }
        __syncthreads();
        /* ---- 1. stage the chunk's returns in shared memory (coalesced loads) and count them into (piece, granule) */
        bool plain = true;
        for (uint32_t e = tid; e < n_ret; e += RR_BLOCK) {
            const uint32_t v = c0 + (rpw == 2u ? (e >> 1) : e), q = (rpw == 2u) ? (e & 1u) : 0u;
            int p = 0;
            while (p + 1 < n_passes && v >= s_voff[p + 1]) p++;
            const size_t idx = 2 * (size_t)(s_begin[p] + (v - s_voff[p])) + q;
            const int cell = __ldcs(reinterpret_cast<const int*>(P.sig_cell + (size_t)p * P.wave_cap) + idx);
            int g_lo, g_hi;
            uint32_t sv = 0u;
            if (window(cell, g_lo, g_hi)) {
                sv = __ldcs(reinterpret_cast<const uint32_t*>(P.sig_strength + (size_t)p * P.wave_cap) + idx);
                const float svf = __uint_as_float(sv);
                plain = plain && (svf >= 0.0f) && (svf < INFINITY);
                const uint32_t row = (e / plen) * (uint32_t)G1;
                for (int g = g_lo; g <= g_hi; g++) {
                    const uint32_t i16 = row + (uint32_t)g;
                    atomicAdd(reinterpret_cast<uint32_t*>(s_tab) + (i16 >> 1), 1u << (16u * (i16 & 1u)));
                }
            }
            s_ret[e] = make_uint2((uint32_t)cell, sv);
        }
        if (!plain) s_plain = 0u;
        __syncthreads();
        /* ---- 2. scan: per granule the prefix over the pieces (a piece's cursor inside the granule's list) and the total ... */
        for (int g = tid; g < n_gran; g += RR_BLOCK) {
            uint32_t run = 0;
#pragma unroll 8
            for (int p = 0; p < RR_DRAW_PIECES; p++) { const uint32_t t = s_tab[p * G1 + g]; s_tab[p * G1 + g] = (uint16_t)run; run += t; }
            s_off[g] = run;
        }
        __syncthreads();
        /* ... then the list offsets over the granules, and the granules that have a list at all */
        if (wid == 0) {
            uint32_t run = 0, nne = 0;
            for (int base = 0; base < n_gran; base += 32) {
                const int g = base + lane;
                const uint32_t tot = (g < n_gran) ? s_off[g] : 0u;
                uint32_t incl = tot;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) { const uint32_t up = __shfl_up_sync(RR_FULL, incl, off); if (lane >= off) incl += up; }
                const uint32_t ne = __ballot_sync(RR_FULL, tot > 0u);
                if (g < n_gran) s_off[g] = run + incl - tot;
                if (tot > 0u) s_ne[nne + __popc(ne & ((1u << lane) - 1u))] = (uint16_t)g;
                nne += __popc(ne);
                run += __shfl_sync(RR_FULL, incl, 31);
            }
            if (lane == 0) { s_off[n_gran] = run; s_nne = nne; }
        }
        __syncthreads();
        /* ---- 3. fill: thread p walks piece p in list order */
        if (tid < RR_DRAW_PIECES) {
            uint16_t* cur = s_tab + tid * G1;
            const uint32_t pb = (uint32_t)tid * plen, pe = min(n_ret, (uint32_t)(tid + 1) * plen);
            int next_cell = (pb < pe) ? (int)s_ret[pb].x : 0;      /* the next return's cell is loaded before this one's stores */
            for (uint32_t e = pb; e < pe; e++) {
                const int cell = next_cell;
                if (e + 1 < pe) next_cell = (int)s_ret[e + 1].x;
                int g_lo, g_hi;
                if (!window(cell, g_lo, g_hi)) continue;
                if (g_hi - g_lo <= 2) {                            /* up to 3 granules (W <= 65): the three cursor chains side by side */
                    const bool h1 = g_hi > g_lo, h2 = g_hi > g_lo + 1;
                    const uint32_t c0 = cur[g_lo], c1 = h1 ? cur[g_lo + 1] : 0u, c2 = h2 ? cur[g_lo + 2] : 0u;
                    const uint32_t o0 = s_off[g_lo], o1 = h1 ? s_off[g_lo + 1] : 0u, o2 = h2 ? s_off[g_lo + 2] : 0u;
                    cur[g_lo] = (uint16_t)(c0 + 1u); s_ent[o0 + c0] = (uint16_t)e;
                    if (h1) { cur[g_lo + 1] = (uint16_t)(c1 + 1u); s_ent[o1 + c1] = (uint16_t)e; }
                    if (h2) { cur[g_lo + 2] = (uint16_t)(c2 + 1u); s_ent[o2 + c2] = (uint16_t)e; }
                } else {
                    for (int g = g_lo; g <= g_hi; g++) { const uint32_t c = cur[g]; cur[g] = (uint16_t)(c + 1u); s_ent[s_off[g] + c] = (uint16_t)e; }
                }
            }
        }
        __syncthreads();
        /* ---- 4. add: lane <-> bin, value in a register, list front to back */
        const uint32_t nne = s_nne;
        const bool plain_chunk = s_plain != 0u;
        for (;;) {
            uint32_t qi = 0;
            if (lane == 0) qi = atomicAdd(&s_next, 1u);
            qi = __shfl_sync(RR_FULL, qi, 0);
            if (qi >= nne) break;
            const uint32_t g = s_ne[qi];
            const uint32_t eb = s_off[g], ee = (g + 1 < (uint32_t)n_gran) ? s_off[g + 1] : s_off[n_gran];
            const int bin = (int)(g << 5) + lane;
            const bool live = (bin >= lo_bin) && (bin < C);
            /* k = bin - (cell - mode); a lane without a bin gets a k that is outside every window */
            const int bin_m = live ? bin + mode : -(1 << 24);
            float acc = live ? s_col[bin] : 0.0f;
            if (P.denoise_on && plain_chunk) {
                /* every strength and weight of the chunk is finite and >= 0: a bin outside a return's window may add the
                 * product with a zero weight instead of being masked ((double)acc + 0.0 == acc exactly, acc is never -0),
                 * and the column only grows, so its running maximum is its last value */
#pragma unroll 4
                for (uint32_t e = eb; e < ee; e++) {
                    const uint2 rt = s_ret[s_ent[e]];
                    const uint32_t k = (uint32_t)(bin_m - (int)rt.x);
                    acc = (float)((double)acc + (double)__uint_as_float(rt.y) * s_weights[min(k, (uint32_t)(RR_MAX_DENOISE - 1))]);
                }
                m = fmaxf(m, acc);
            } else if (P.denoise_on) {
                const uint32_t wmax = (uint32_t)W - 1u;
#pragma unroll 4
                for (uint32_t e = eb; e < ee; e++) {
                    const uint2 rt = s_ret[s_ent[e]];
                    const uint32_t k = (uint32_t)(bin_m - (int)rt.x);
                    const float v = (float)((double)acc + (double)__uint_as_float(rt.y) * s_weights[min(k, wmax)]);
                    acc = (k <= wmax) ? v : acc;                   /* bins outside the window keep their value */
                    m = fmaxf(m, acc);                             /* == if (acc > m) m = acc, NaN never enters (:428-431) */
                }
            } else {
                for (uint32_t e = eb; e < ee; e++) {
                    const uint2 rt = s_ret[s_ent[e]];
                    const float sv = __uint_as_float(rt.y);
                    const float v = (acc < sv) ? sv : acc;         /* std::max(old, strength), :439 */
                    acc = (live && bin == (int)rt.x) ? v : acc;
                    m = fmaxf(m, acc);
                }
            }
            if (live) s_col[bin] = acc;
        }
        __syncthreads();
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const float o2 = __shfl_xor_sync(RR_FULL, m, off); if (o2 > m) m = o2; }
    if (lane == 0) s_red[wid] = m;
    __syncthreads();
    float max_val = 0.0f;
#pragma unroll
    for (int k = 0; k < RR_WARPS; k++) if (s_red[k] > max_val) max_val = s_red[k];

    /* ================= energy_max, ambient noise, normalise, mono8 (RadarCPU.cpp:453-542) ================= */
    const int col = (((P.scroll_image + az) % RR_N_ANGLES) + RR_N_ANGLES) % RR_N_ANGLES;   /* RadarCPU.cpp:457; scroll_image is validated to [0, 400] */
    const uint64_t frame_id = P.frame_id0 + (uint64_t)pose_i;
    const float signal_amp = max_val - 0.0f;
    const float noise_at_0 = (float)((double)signal_amp * P.noise_at_signal_0);
    const float noise_at_1 = (float)((double)signal_amp * P.noise_at_signal_1);
    const float ne_max = (float)((double)max_val * P.noise_energy_max);
    const float ne_min = (float)((double)max_val * P.noise_energy_min);
    const double random_begin = (P.ambient_noise == 2)
        ? (double)rr_noise_u01(P.noise_seed, frame_id, (uint32_t)az, 0u) * 1000.0 : 0.0;
    const float out_scale = (float)(P.signal_max / (double)max_val);
    const RRPerlinRow row1 = rr_perlin_row((double)col * 0.05), row2 = rr_perlin_row((double)col * 0.2);
    uint8_t* out_rows = P.out + (size_t)pose_i * (size_t)C * RR_N_ANGLES + col;
    for (int i = tid; i < C; i += RR_BLOCK) {
        float v = s_col[i] * P.energy_max_f;
        if (P.ambient_noise) {
            double p = 0.0;
            if (P.ambient_noise == 1) {
                p = (double)rr_noise_u01(P.noise_seed, frame_id, (uint32_t)az, 1u + (uint32_t)i);
            } else if (P.ambient_noise == 2) {
                const double p1 = rr_perlin2(s_perm, s_grad, random_begin + (double)i * 0.05, row1);
                const double p2 = rr_perlin2(s_perm, s_grad, random_begin + (double)i * 0.2, row2);
                p = 0.9 * p1 + 0.1 * p2;
            }
            const float sn = (float)(1.0 - (double)((v - 0.0f) / signal_amp));
            const float sn4 = (float)rr_pow4(sn);
            const float amp = (float)((double)(sn4 * noise_at_0) + (1.0 - (double)sn4) * (double)noise_at_1);
            float y = (float)((double)amp * p);
            y = y + (ne_max - ne_min) * __ldg(P.noise_decay + i) + ne_min;   /* exp(-loss * x_i), :517-521 */
            y = fabsf(y);
            v = v + y;
        }
        v = v * out_scale;
        if (DEBUG && P.dbg_columns) P.dbg_columns[(size_t)az * C + i] = v;
        const uint8_t px = rr_to_u8(v);
        if (OUT == RR_OUT_BYTES) out_rows[(size_t)i * RR_N_ANGLES] = px;
        else s_bytes[i] = px;
    }
    if (OUT != RR_OUT_BYTES && tid < C16 - C) s_bytes[C + tid] = 0;      /* padding of the staged column */
    /* 8 adjacent azimuths -> 8-byte row segments. For both variants the launcher guarantees: az_count % 8 == 0 (a group
     * never straddles two poses), (scroll_image + az_begin) % 8 == 0 (segments are aligned and never wrap at column
     * 400) and an 8-byte aligned image. */
    if (OUT == RR_OUT_GROUP) {
        __syncthreads();
        const uint32_t r = item % RR_DRAW_GROUP, grp = item / RR_DRAW_GROUP;
        uint4* mine = reinterpret_cast<uint4*>(P.draw_stage + (size_t)item * (size_t)C16);
        for (int k = tid; k < (C16 >> 4); k += RR_BLOCK) __stcg(mine + k, reinterpret_cast<const uint4*>(s_bytes)[k]);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(P.draw_group_done + grp, 1u) == RR_DRAW_GROUP - 1) ? 1u : 0u;
        __syncthreads();
        if (s_last) {                                               /* the other 7 columns are complete and visible */
            __threadfence();
            const uint32_t* src = reinterpret_cast<const uint32_t*>(P.draw_stage + (size_t)grp * RR_DRAW_GROUP * (size_t)C16);
            uint8_t* seg = out_rows - r;                            /* column of the group's first azimuth */
            const int n_words = (C + 3) >> 2, stride = C16 >> 2;
            for (int w = tid; w < n_words; w += RR_BLOCK) {
                uint32_t x[RR_DRAW_GROUP];
#pragma unroll
                for (int j = 0; j < RR_DRAW_GROUP; j++) x[j] = __ldcg(src + (size_t)j * stride + w);
                rr_store_segments(x, seg, 4 * w, C);
            }
        }
    }
    if (OUT == RR_OUT_CLUSTER) {
        namespace cg = cooperative_groups;
        cg::cluster_group cl = cg::this_cluster();
        cl.sync();
        const unsigned r = cl.block_rank();
        const uint32_t* rem[RR_DRAW_GROUP];
#pragma unroll
        for (int j = 0; j < RR_DRAW_GROUP; j++) rem[j] = reinterpret_cast<const uint32_t*>(cl.map_shared_rank(s_bytes, j));
        const int n_words = (C + 3) >> 2, wpr = (n_words + RR_DRAW_GROUP - 1) / RR_DRAW_GROUP;
        const int w_end = min(n_words, (int)(r + 1) * wpr);
        uint8_t* seg = out_rows - r;
        for (int w = (int)r * wpr + tid; w < w_end; w += RR_BLOCK) {
            uint32_t x[RR_DRAW_GROUP];
#pragma unroll
            for (int j = 0; j < RR_DRAW_GROUP; j++) x[j] = rem[j][w];
            rr_store_segments(x, seg, 4 * w, C);
        }
        cl.sync();                                                  /* nobody leaves while its column is still being read */
    }
    if (OUT == RR_OUT_COLUMNS) {
        __syncthreads();
        auto copy_column = [&](uint8_t* dst) {
            if (((C & 15) == 0) && ((reinterpret_cast<size_t>(dst) & 15) == 0)) {
                const uint4* src4 = reinterpret_cast<const uint4*>(s_bytes);
                uint4* dst4 = reinterpret_cast<uint4*>(dst);
                for (int k = tid; k < (C >> 4); k += RR_BLOCK) dst4[k] = src4[k];
            } else {
                for (int k = tid; k < C; k += RR_BLOCK) dst[k] = s_bytes[k];
            }
        };
        if (P.n_peers > 0) {
            /* the finished mono8 column goes to every rank's gather buffer with 16-byte peer stores (NVLink) */
            const size_t col_off = ((size_t)(P.peer_pose0 + (uint32_t)pose_i) * RR_N_ANGLES + (size_t)az) * (size_t)C;
            for (int p = 0; p < P.n_peers; p++) copy_column(P.peer_out[p] + col_off);
        } else {
            copy_column(P.out + ((size_t)pose_i * P.az_count + (size_t)(az - P.az_begin)) * (size_t)C);
        }
    }
}

/* Sum of squared pixel differences of every rendered goal image against a recorded ("real") polar image: the data
 * term of the material optimiser's objective, -PSNR (scripts/radaray_opti.py:198, skimage peak_signal_noise_ratio =
 * 10 log10(255^2 / mean((a-b)^2))). Exact integer arithmetic; grid = (chunks, goals). */
__global__ void __launch_bounds__(256) rr_score_kernel(const uint8_t* __restrict__ sim, const uint8_t* __restrict__ real,
                                                       size_t img_bytes, size_t real_stride, unsigned long long* ssd)
{
    const size_t goal = blockIdx.y;
    const uint4* a = reinterpret_cast<const uint4*>(sim + goal * img_bytes);
    const uint4* b = reinterpret_cast<const uint4*>(real + goal * real_stride);
    const size_t n16 = img_bytes / 16;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = a[i], y = __ldg(b + i);
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
        uint32_t part = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int sft = 0; sft < 32; sft += 8) {
                const int d = (int)((xs[k] >> sft) & 0xffu) - (int)((ys[k] >> sft) & 0xffu);
                part += (uint32_t)(d * d);
            }
        }
        acc += part;
    }
    for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < img_bytes; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)sim[goal * img_bytes + i] - (int)real[goal * real_stride + i];
        acc += (unsigned long long)(d * d);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(RR_FULL, acc, off);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(ssd + goal, acc);
}

extern "C" cudaError_t rr_launch_score(const uint8_t* sim, const uint8_t* real, size_t img_bytes, size_t real_stride,
                                       size_t n_goals, unsigned long long* ssd, cudaStream_t st)
{
    if (!n_goals) return cudaSuccess;
    const unsigned chunks = (unsigned)std::min<size_t>(64, (img_bytes / 16 + 255) / 256 + 1);
    rr_score_kernel<<<dim3(chunks, (unsigned)n_goals), 256, 0, st>>>(sim, real, img_bytes, real_stride, ssd);
    return cudaGetLastError();
}

/* ---- azimuth-sharded frames: completion flags over peer memory + local transpose ----------------------------------
 * rank r, after its draw kernel (whose peer stores are complete when the kernel ends), publishes `epoch` in slot r of
 * every rank's flag array; a rank may read its gather buffer once all `world` slots of its own array show the epoch. */
struct RRPeerFlags { uint32_t* flags[RR_MAX_PEERS]; };

__global__ void rr_peer_signal_kernel(const RRPeerFlags F, int rank, int world, uint32_t epoch)
{
    const int p = threadIdx.x;
    if (p >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(F.flags[p] + rank), "r"(epoch) : "memory");
}

__global__ void rr_peer_wait_kernel(const uint32_t* my_flags, int world, uint32_t epoch, int32_t* error_flags,
                                    volatile int32_t* host_sticky, unsigned long long timeout_ns)
{
    const int p = threadIdx.x;
    if (p >= world) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + p) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {                                /* a peer never arrived: the frame below is incomplete */
            atomicExch(&error_flags[2], 1);
            *host_sticky = p + 1;                                  /* zero-copy host flag: stays set until the caller has seen it */
            __threadfence_system();
            break;
        }
        __nanosleep(200);
    }
}

/* gather buffer [pose][400][C] (column-major) -> the reference's image [pose][C][400] with the scroll applied */
__global__ void __launch_bounds__(256) rr_gather_transpose_kernel(const uint8_t* __restrict__ gather, uint8_t* __restrict__ out,
                                                                  int C, int scroll)
{
    __shared__ uint8_t tile[32][33];
    const size_t pose = blockIdx.z;
    const int a0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          /* 32 x 8 threads */
    const uint8_t* g = gather + pose * (size_t)RR_N_ANGLES * C;
    uint8_t* o = out + pose * (size_t)C * RR_N_ANGLES;
    for (int r = ty; r < 32; r += 8) {                                /* rows = azimuths, columns = cells (contiguous) */
        const int a = a0 + r, c = c0 + tx;
        tile[r][tx] = (a < RR_N_ANGLES && c < C) ? g[(size_t)a * C + c] : (uint8_t)0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {                                /* rows = cells, columns = azimuths (contiguous) */
        const int c = c0 + r, a = a0 + tx;
        if (c < C && a < RR_N_ANGLES) o[(size_t)c * RR_N_ANGLES + (((scroll + a) % RR_N_ANGLES) + RR_N_ANGLES) % RR_N_ANGLES] = tile[tx][r];
    }
}

extern "C" cudaError_t rr_launch_peer_exchange(uint32_t* const* peer_flags, int rank, int world, uint32_t epoch,
                                               const uint8_t* my_gather, uint8_t* d_out, int n_cells, int scroll, int n_poses,
                                               int32_t* error_flags, int32_t* host_sticky, unsigned long long timeout_ns, cudaStream_t st)
{
    RRPeerFlags F;
    for (int p = 0; p < RR_MAX_PEERS; p++) F.flags[p] = p < world ? peer_flags[p] : nullptr;
    rr_peer_signal_kernel<<<1, 32, 0, st>>>(F, rank, world, epoch);
    rr_peer_wait_kernel<<<1, 32, 0, st>>>(peer_flags[rank], world, epoch, error_flags, host_sticky, timeout_ns);
    const dim3 grid((unsigned)((n_cells + 31) / 32), (unsigned)((RR_N_ANGLES + 31) / 32), (unsigned)n_poses);
    rr_gather_transpose_kernel<<<grid, 256, 0, st>>>(my_gather, d_out, n_cells, scroll);
    return cudaGetLastError();
}

/* raw closest-hit probe (rr_cast_rays) */
__global__ void rr_cast_kernel(const RRNode* nodes, const float4* tris, uint32_t root_ref,
                               float gox, float goy, float goz, float gsx, float gsy, float gsz,
                               const float* origins, const float* dirs, size_t n, float tmax,
                               int32_t* face_ids, float* ranges)
{
    extern __shared__ uint32_t s_cast_stack[];
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float go[3] = {gox, goy, goz}, gs[3] = {gsx, gsy, gsz};
    unsigned a = 0, b = 0;
    float t; int face;
    rr_trace<false>(nodes, tris, root_ref, go, gs, rr_v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
                    rr_v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), tmax, t, face, a, b, s_cast_stack, __activemask());
    face_ids[i] = face;
    ranges[i] = t;
}

/* ---- host-side launchers (called from rr_api.cu) ------------------------------------------------*/
extern "C" cudaError_t rr_launch_trace(const RRFrameParams* P, int pass, int grid, cudaStream_t st, int stats, int debug)
{
    if (debug) rr_trace_kernel<true, true><<<grid, RR_TRACE_BLOCK, RR_TRACE_SMEM_BYTES, st>>>(*P, pass);
    else if (stats) rr_trace_kernel<true, false><<<grid, RR_TRACE_BLOCK, RR_TRACE_SMEM_BYTES, st>>>(*P, pass);
    else rr_trace_kernel<false, false><<<grid, RR_TRACE_BLOCK, RR_TRACE_SMEM_BYTES, st>>>(*P, pass);
    return cudaGetLastError();
}

extern "C" cudaError_t rr_launch_prep(const RRFrameParams* P, cudaStream_t st)
{
    rr_prep_kernel<<<(unsigned)((P->n_items + 127) / 128), 128, 0, st>>>(*P);
    return cudaGetLastError();
}

extern "C" cudaError_t rr_launch_mat_pairs(const float4* materials, int n_mat, int n_tables, RRMatPair* out, cudaStream_t st)
{
    const int n = n_tables * (n_mat + 1);
    if (n <= 0) return cudaSuccess;
    rr_mat_pairs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(materials, n_mat, n_tables, out);
    return cudaGetLastError();
}

extern "C" cudaError_t rr_launch_scan(const RRFrameParams* P, int pass, cudaStream_t st)
{
    /* one CTA per 1024 groups / 1024 items that the list can hold; CTAs beyond the actual list length exit at once */
    const uint32_t n_super = (P->group_cap + RR_SCAN_BLOCK - 1) / RR_SCAN_BLOCK, n_isuper = ((uint32_t)P->n_items + RR_SCAN_BLOCK - 1) / RR_SCAN_BLOCK;
    const uint32_t grid = n_super > n_isuper ? n_super : n_isuper;
    rr_scan_kernel<<<grid ? grid : 1, RR_SCAN_BLOCK, 0, st>>>(*P, pass);
    return cudaGetLastError();
}

extern "C" size_t rr_draw_smem_bytes(int n_cells)
{
    const size_t c16 = (size_t)((n_cells + 15) & ~15), g1 = (size_t)((((n_cells + 31) >> 5) + 2) & ~1);
    const size_t lists = (size_t)RR_DRAW_RETURNS * 8 + (size_t)RR_DRAW_ENTRIES * 2;
    return c16 * 4 + std::max(lists, c16) + (size_t)RR_DRAW_PIECES * g1 * 2 + g1 * 4 + g1 * 2;
}

template <bool DEBUG, int OUT>
static cudaError_t rr_draw_launch_one(const RRFrameParams& P, int n_items, size_t smem, cudaStream_t st)
{
    /* more than 48 KB of dynamic shared memory is an opt-in per function AND per device: cheap enough to repeat per launch */
    cudaError_t e = cudaFuncSetAttribute(rr_draw_kernel<DEBUG, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024));
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_items); cfg.blockDim = dim3(RR_BLOCK); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (OUT == RR_OUT_CLUSTER) {
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = RR_DRAW_GROUP; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, rr_draw_kernel<DEBUG, OUT>, P);
}

extern "C" cudaError_t rr_launch_draw(const RRFrameParams* P, int n_items, cudaStream_t st, int debug)
{
    const size_t smem = rr_draw_smem_bytes(P->n_cells);
    if (P->n_peers > 0 || P->column_major) return rr_draw_launch_one<false, RR_OUT_COLUMNS>(*P, n_items, smem, st);
    if (debug) return rr_draw_launch_one<true, RR_OUT_BYTES>(*P, n_items, smem, st);
    /* row-major image: 8-byte row segments over groups of 8 adjacent azimuths when the columns of a group are aligned */
    static const int use_cluster = getenv("RR_DRAW_USE_CLUSTER") ? atoi(getenv("RR_DRAW_USE_CLUSTER")) : 0;
    static const int use_bytes = getenv("RR_DRAW_BYTES") ? atoi(getenv("RR_DRAW_BYTES")) : 0;
    const bool aligned = (P->az_count % RR_DRAW_GROUP == 0) && ((P->scroll_image + P->az_begin) % RR_DRAW_GROUP == 0)
                         && ((reinterpret_cast<size_t>(P->out) & 7) == 0) && !use_bytes;
    if (!aligned) return rr_draw_launch_one<false, RR_OUT_BYTES>(*P, n_items, smem, st);
    if (use_cluster) return rr_draw_launch_one<false, RR_OUT_CLUSTER>(*P, n_items, smem, st);
    return rr_draw_launch_one<false, RR_OUT_GROUP>(*P, n_items, smem, st);
}

extern "C" cudaError_t rr_trace_occupancy(int* blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, rr_trace_kernel<false, false>, RR_TRACE_BLOCK, RR_TRACE_SMEM_BYTES);
}

/* the pass as two kernels (rr_pass_body): resident CTAs per SM of each */
extern "C" cudaError_t rr_split_occupancy(int* walk_blocks_per_sm, int* shade_blocks_per_sm)
{
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(walk_blocks_per_sm, rr_walk_kernel<false>, RR_TRACE_BLOCK, RR_TRACE_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(shade_blocks_per_sm, rr_shade_kernel<false>, RR_TRACE_BLOCK, 0);
}

extern "C" cudaError_t rr_dual_occupancy(int* blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, rr_dual_kernel<false, false>, RR_TRACE_BLOCK, 0);
}

extern "C" cudaError_t rr_launch_dual(const RRFrameParams* P, int pass, int grid, cudaStream_t st, int stats, int debug)
{
    if (debug) rr_dual_kernel<true, true><<<grid, RR_TRACE_BLOCK, 0, st>>>(*P, pass);
    else if (stats) rr_dual_kernel<true, false><<<grid, RR_TRACE_BLOCK, 0, st>>>(*P, pass);
    else rr_dual_kernel<false, false><<<grid, RR_TRACE_BLOCK, 0, st>>>(*P, pass);
    return cudaGetLastError();
}

extern "C" cudaError_t rr_launch_walk(const RRFrameParams* P, int pass, int grid, cudaStream_t st, int stats)
{
    if (stats) rr_walk_kernel<true><<<grid, RR_TRACE_BLOCK, RR_TRACE_SMEM_BYTES, st>>>(*P, pass);
    else rr_walk_kernel<false><<<grid, RR_TRACE_BLOCK, RR_TRACE_SMEM_BYTES, st>>>(*P, pass);
    return cudaGetLastError();
}

extern "C" cudaError_t rr_launch_shade(const RRFrameParams* P, int pass, int grid, cudaStream_t st, int debug)
{
    if (debug) rr_shade_kernel<true><<<grid, RR_TRACE_BLOCK, 0, st>>>(*P, pass);
    else rr_shade_kernel<false><<<grid, RR_TRACE_BLOCK, 0, st>>>(*P, pass);
    return cudaGetLastError();
}

extern "C" cudaError_t rr_launch_cast(const RRNode* nodes, const float4* tris, uint32_t root_ref, const float* go,
                                      const float* gs, const float* origins, const float* dirs, size_t n, float tmax,
                                      int32_t* face_ids, float* ranges, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const int block = RR_TRACE_BLOCK;                            /* the shared-memory stack is laid out for this block size */
    const unsigned grid = (unsigned)((n + block - 1) / block);
    rr_cast_kernel<<<grid, block, RR_TRACE_SMEM_BYTES, st>>>(nodes, tris, root_ref, go[0], go[1], go[2], gs[0], gs[1], gs[2],
                                           origins, dirs, n, tmax, face_ids, ranges);
    return cudaGetLastError();
}
