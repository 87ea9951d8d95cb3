// integrate.cuh — the functor kernels of integrate() (integrate_nh.h:29-55) as streaming CUDA kernels, the Philox
// counter-based generator that stands in for rng.h's per-thread MT19937 + xorshift128 stream, the temperature reduction
// and the volume constraint.
//
// Replaces integrate_nh.h:58-273, integrate_langevin.h:99-149, openrbc.cpp:114-131, compute_temperature.h:23-29,
// constrain_volume.h:26-83, rng.h / math_vector_integer.h:62-66 (semantics: r uniform in [-1,1) with 2^-31 grain).
#pragma once
#include "common.cuh"
#include "pair.cuh"

namespace orbc {

// ---- Philox4x32-10 (Salmon et al., SC'11); counter = (particle index, step, species, 0), key = 64-bit seed -----------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    #pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// math_vector_integer.h:62-66: u * 2^-31 - 1 (one multiply, one subtract, both rounded)
__device__ __forceinline__ float uint2u11(uint32_t u) { return __fsub_rn(__fmul_rn(__uint2float_rn(u), 4.6566129e-10f), 1.0f); }

__device__ __forceinline__ F3 noise3(uint64_t seed, uint32_t step, uint32_t species, uint32_t i) {
    uint32_t o[4];
    philox4x32_10(i, step, species, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return {uint2u11(o[0]), uint2u11(o[1]), uint2u11(o[2])};
}

__global__ void k_noise(uint64_t seed, uint32_t step, uint32_t species, size_t n, float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F3 r = noise3(seed, step, species, (uint32_t)i);
    out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
}

// Halo push of a decomposed run, fused into the integrator: the new position and director of an owned particle are also
// stored straight into the arrays of the ranks that read it as halo (same global slot, peer memory over NVLink).
struct PushArgs {
    const unsigned char *cell_mask;        // per cell: ranks owning a member of its r<9 stencil (rebuild.cuh HaloOut)
    const unsigned char *pmask;            // per protein slot: ranks owning a bonded partner (null for lipids)
    const int *cellid;
    float4 *x[kMaxWorld], *nn[kMaxWorld];
    int world;
};
struct IntegArgs {
    float4 *x, *v, *f, *nn, *o, *t;
    float4 *x_out, *nn_out;                // where verlet_langevin stores the new x, n (the same arrays on a single GPU)
    size_t n;
    const int *range;                      // {begin, end} slots to integrate (verlet_langevin)
    int clear;                             // verlet_langevin: zero f and t (the reference's semantics) or leave them dead
    PushArgs push;
    int species;
    float dt;
    float gamma[kNType], sigma[kNType];   // Langevin: 6 pi eta R, sqrt(2 kBT gamma) sqrt(3/dt)   (integrate_langevin.h:110-114)
    float zeta;
    float lo, hi; double dlo, dhi;        // box (runtime_parameter.h:111-115)
    double dt_d, dr_opt, dn_opt;
    uint64_t seed; uint32_t step;
    const float *noise;                    // optional injected noise (n x 3), device pointer
    double *acc;                           // acc[0] += kinetic energy
    const float *zeta_dev;                 // if non-null, zeta is read from the device (orbc_run_nh)
    unsigned *disp;                        // if non-null: largest squared displacement of this step, for the hit lists (NlState::disp)
};

__device__ __forceinline__ F3 cross3(F3 u, F3 v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
__device__ __forceinline__ F3 rotate_director(F3 n, F3 o, float dt) {
    // n = normalize(n + cross(o, n) * dt)    integrate_langevin.h:126 / integrate_nh.h:218
    const F3 c = cross3(o, n);
    F3 g = {n.x + c.x * dt, n.y + c.y * dt, n.z + c.z * dt};
    const float s = 1.0f / sqrtf(g.x * g.x + g.y * g.y + g.z * g.z);
    return {g.x * s, g.y * s, g.z * s};
}
__device__ __forceinline__ void bounce(float &x, float &v, double lo, double hi) {
    // integrate_nh.h:200-209
    if (x < lo) { x = (float)(lo + (lo - x)); v = -v; }
    else if (x > hi) { x = (float)(hi - (x - hi)); v = -v; }
}

__device__ __forceinline__ void block_add_double(double v, double *dst) {
    __shared__ double s[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) s[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (blockDim.x >> 5) ? s[lane] : 0.0;
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) atomicAdd(dst, v);
    }
}

// integrate_nh.h:58-67
__global__ void k_clear_force(float4 *__restrict__ f, float4 *__restrict__ t, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f[i] = make_float4(0, 0, 0, 0); t[i] = make_float4(0, 0, 0, 0);
}
// integrate_nh.h:146-154
__global__ void k_post_torque(const float4 *__restrict__ nn, float4 *__restrict__ t, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = nn[i], b = t[i];
    const F3 r = cross3({a.x, a.y, a.z}, {b.x, b.y, b.z});
    t[i] = make_float4(r.x, r.y, r.z, 0);
}
// integrate_nh.h:124-144
__global__ void k_bounce_back(IntegArgs a) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float4 x = a.x[i], v = a.v[i];
    bounce(x.x, v.x, a.dlo, a.dhi); bounce(x.y, v.y, a.dlo, a.dhi); bounce(x.z, v.z, a.dlo, a.dhi);
    a.x[i] = x; a.v[i] = v;
}

// integrate_langevin.h:99-149 — one pass: torque -> omega -> director, noise, friction, kick, drift, clear f and t
__device__ __forceinline__ void verlet_langevin_body(const unsigned bid, const IntegArgs &a) {
    const size_t i = (size_t)a.range[0] + (size_t)bid * blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (i < (size_t)a.range[1]) {
        float4 x = a.x[i], v = a.v[i], n4 = a.nn[i], o = a.o[i];
        const float4 f = a.f[i], t = a.t[i];
        const int type = __float_as_int(x.w);
        const F3 tq = cross3({n4.x, n4.y, n4.z}, {t.x, t.y, t.z});
        o.x += a.dt * tq.x; o.y += a.dt * tq.y; o.z += a.dt * tq.z;
        const F3 nn = rotate_director({n4.x, n4.y, n4.z}, {o.x, o.y, o.z}, a.dt);
        F3 r = {0.f, 0.f, 0.f};
        if (a.noise) r = {a.noise[3 * i], a.noise[3 * i + 1], a.noise[3 * i + 2]};
        else if (a.sigma[type] != 0.f) r = noise3(a.seed, a.step, a.species, (uint32_t)i);
        const float g = a.gamma[type], s = a.sigma[type];
        const float fx = f.x - (g * v.x + s * r.x), fy = f.y - (g * v.y + s * r.y), fz = f.z - (g * v.z + s * r.z);
        const float k = a.dt / c_ff.mass[type];
        v.x += fx * k; v.y += fy * k; v.z += fz * k;
        x.x += v.x * a.dt; x.y += v.y * a.dt; x.z += v.z * a.dt;
        d2 = (v.x * v.x + v.y * v.y + v.z * v.z) * (a.dt * a.dt);
        a.x_out[i] = x; a.v[i] = v; a.o[i] = o;
        const float4 nnew = make_float4(nn.x, nn.y, nn.z, n4.w);
        a.nn_out[i] = nnew;
        if (a.clear) { a.f[i] = make_float4(0, 0, 0, 0); a.t[i] = make_float4(0, 0, 0, 0); }
        if (a.push.world > 1) {
            unsigned m = a.push.cell_mask[a.push.cellid[i]];
            if (a.push.pmask) m |= a.push.pmask[i];
            while (m) {
                const int r = __ffs(m) - 1; m &= m - 1;
                a.push.x[r][i] = x; a.push.nn[r][i] = nnew;
            }
        }
    }
    nl_track(a.disp, d2);
}
__global__ void __launch_bounds__(256) k_verlet_langevin(IntegArgs a) { verlet_langevin_body(blockIdx.x, a); }
// both containers in one launch: blocks [0, blocks0) integrate a0's particles, the others a1's
__global__ void __launch_bounds__(256) k_verlet_langevin2(IntegArgs a0, IntegArgs a1, unsigned blocks0) {
    if (blockIdx.x < blocks0) verlet_langevin_body(blockIdx.x, a0); else verlet_langevin_body(blockIdx.x - blocks0, a1);
}

// integrate_nh.h:178-235 (operator()) — half kick with 1/(1 + dt zeta / 2), drift, bounce-back, KE, omega half kick, director, clear
__device__ __forceinline__ void nh_initial_fused_body(const unsigned bid, const IntegArgs &a) {
    const size_t i = (size_t)a.range[0] + (size_t)bid * blockDim.x + threadIdx.x;
    double ke = 0.0;
    float d2 = 0.f;
    if (i < (size_t)a.range[1]) {
        const float zeta = a.zeta_dev ? a.zeta_dev[0] : a.zeta;
        const float gamma = 1.0f / (1.0f + 0.5f * a.dt * zeta);
        float4 x = a.x[i], v = a.v[i], n4 = a.nn[i], o = a.o[i];
        const float4 f = a.f[i], t = a.t[i];
        const int type = __float_as_int(x.w);
        const float m = c_ff.mass[type];
        const float s = 0.5f / m * a.dt;
        v.x = (v.x + s * f.x) * gamma; v.y = (v.y + s * f.y) * gamma; v.z = (v.z + s * f.z) * gamma;
        x.x += v.x * a.dt; x.y += v.y * a.dt; x.z += v.z * a.dt;
        bounce(x.x, v.x, a.dlo, a.dhi); bounce(x.y, v.y, a.dlo, a.dhi); bounce(x.z, v.z, a.dlo, a.dhi);   // a reflection never lengthens the step
        d2 = (v.x * v.x + v.y * v.y + v.z * v.z) * (a.dt * a.dt);
        ke = 0.5f * m * (v.x * v.x + v.y * v.y + v.z * v.z);
        const float so = 0.5f * a.dt;
        o.x += so * t.x; o.y += so * t.y; o.z += so * t.z;
        const F3 nn = rotate_director({n4.x, n4.y, n4.z}, {o.x, o.y, o.z}, a.dt);
        const float4 nnew = make_float4(nn.x, nn.y, nn.z, n4.w);
        a.x_out[i] = x; a.v[i] = v; a.o[i] = o;
        a.nn_out[i] = nnew;
        a.f[i] = make_float4(0, 0, 0, 0); a.t[i] = make_float4(0, 0, 0, 0);
        if (a.push.world > 1) {                                  // halo push, as in k_verlet_langevin
            unsigned m = a.push.cell_mask[a.push.cellid[i]];
            if (a.push.pmask) m |= a.push.pmask[i];
            while (m) {
                const int r = __ffs(m) - 1; m &= m - 1;
                a.push.x[r][i] = x; a.push.nn[r][i] = nnew;
            }
        }
    }
    nl_track(a.disp, d2);
    block_add_double(ke, a.acc);
}
__global__ void __launch_bounds__(256) k_nh_initial_fused(IntegArgs a) { nh_initial_fused_body(blockIdx.x, a); }
__global__ void __launch_bounds__(256) k_nh_initial_fused2(IntegArgs a0, IntegArgs a1, unsigned blocks0) {   // both containers, as k_verlet_langevin2
    if (blockIdx.x < blocks0) nh_initial_fused_body(blockIdx.x, a0); else nh_initial_fused_body(blockIdx.x - blocks0, a1);
}

// integrate_nh.h:237-273 — t = n x t, second half kick with -zeta v, omega half kick, KE
__device__ __forceinline__ void nh_final_fused_body(const unsigned bid, const IntegArgs &a) {
    const size_t i = (size_t)a.range[0] + (size_t)bid * blockDim.x + threadIdx.x;
    double ke = 0.0;
    if (i < (size_t)a.range[1]) {
        const float zeta = a.zeta_dev ? a.zeta_dev[0] : a.zeta;
        float4 v = a.v[i], o = a.o[i];
        const float4 x = a.x[i], n4 = a.nn[i], f = a.f[i], t = a.t[i];
        const int type = __float_as_int(x.w);
        const float m = c_ff.mass[type];
        const F3 tq = cross3({n4.x, n4.y, n4.z}, {t.x, t.y, t.z});
        const float s = 0.5f * a.dt;
        v.x += s * (f.x / m - zeta * v.x); v.y += s * (f.y / m - zeta * v.y); v.z += s * (f.z / m - zeta * v.z);
        o.x += s * tq.x; o.y += s * tq.y; o.z += s * tq.z;
        ke = 0.5f * m * (v.x * v.x + v.y * v.y + v.z * v.z);
        a.v[i] = v; a.o[i] = o; a.t[i] = make_float4(tq.x, tq.y, tq.z, 0);
    }
    block_add_double(ke, a.acc);
}
__global__ void __launch_bounds__(256) k_nh_final_fused(IntegArgs a) { nh_final_fused_body(blockIdx.x, a); }
__global__ void __launch_bounds__(256) k_nh_final_fused2(IntegArgs a0, IntegArgs a1, unsigned blocks0) {
    if (blockIdx.x < blocks0) nh_final_fused_body(blockIdx.x, a0); else nh_final_fused_body(blockIdx.x - blocks0, a1);
}

// integrate_nh.h:156-176 (unfused second half: no torque conversion, no KE)
__global__ void k_nh_final(IntegArgs a) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float4 v = a.v[i], o = a.o[i];
    const float4 x = a.x[i], f = a.f[i], t = a.t[i];
    const float m = c_ff.mass[__float_as_int(x.w)];
    const float s = 0.5f * a.dt;
    v.x += s * (f.x / m - a.zeta * v.x); v.y += s * (f.y / m - a.zeta * v.y); v.z += s * (f.z / m - a.zeta * v.z);
    o.x += s * t.x; o.y += s * t.y; o.z += s * t.z;
    a.v[i] = v; a.o[i] = o;
}

// integrate_nh.h:69-94 (KE only) and compute_temperature.h:23-29 (sum m v^2): acc[0] += scale * m |v|^2
__global__ void __launch_bounds__(256) k_kinetic(const float4 *__restrict__ x, const float4 *__restrict__ v, const int *__restrict__ range, float scale, double *acc) {
    const size_t i = (size_t)range[0] + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < (size_t)range[1]) {
        const float4 vv = v[i];
        e = scale * c_ff.mass[__float_as_int(x[i].w)] * (vv.x * vv.x + vv.y * vv.y + vv.z * vv.z);
    }
    block_add_double(e, acc);
}

// openrbc.cpp:114-131 — capped steepest-descent move of the energy-minimisation loop
__global__ void k_opt_move(IntegArgs a) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float4 x = a.x[i], n4 = a.nn[i];
    const float4 f = a.f[i], t = a.t[i];
    const float m = c_ff.mass[__float_as_int(x.w)];
    const float dxn = sqrtf((f.x / m) * (f.x / m) + (f.y / m) * (f.y / m) + (f.z / m) * (f.z / m));
    const F3 dn = cross3({t.x, t.y, t.z}, {n4.x, n4.y, n4.z});
    const float dnn = sqrtf(dn.x * dn.x + dn.y * dn.y + dn.z * dn.z);
    double dt = a.dt_d;
    if (dxn > a.dr_opt || dnn > a.dn_opt) dt = fmin(a.dr_opt / dxn, a.dn_opt / dnn);
    const float sx = (float)(dt / m), sn = (float)dt;
    x.x += f.x * sx; x.y += f.y * sx; x.z += f.z * sx;
    float gx = n4.x + dn.x * sn, gy = n4.y + dn.y * sn, gz = n4.z + dn.z * sn;
    const float s = 1.0f / sqrtf(gx * gx + gy * gy + gz * gz);
    a.x[i] = x;
    a.nn[i] = make_float4(gx * s, gy * s, gz * s, n4.w);
}

// One iteration of the energy-minimisation loop behind the forces (openrbc.cpp:110-133) as ONE pass: post_torque
// (integrate_nh.h:146-154), the capped steepest-descent mover (openrbc.cpp:114-131, same arithmetic as k_opt_move) and
// bounce_back (integrate_nh.h:124-144); `clear` also does the clear_force of the next iteration (openrbc.cpp:94).  Range-based
// and out of place like k_verlet_langevin, so it also serves a decomposed run (halo push fused in).
__global__ void __launch_bounds__(256) k_opt_fused(IntegArgs a) {
    const size_t i = (size_t)a.range[0] + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)a.range[1]) return;
    float4 x = a.x[i], n4 = a.nn[i];
    const float4 f = a.f[i], t0 = a.t[i];
    const F3 t = cross3({n4.x, n4.y, n4.z}, {t0.x, t0.y, t0.z});                 // post_torque
    const float m = c_ff.mass[__float_as_int(x.w)];
    const float dxn = sqrtf((f.x / m) * (f.x / m) + (f.y / m) * (f.y / m) + (f.z / m) * (f.z / m));
    const F3 dn = cross3(t, {n4.x, n4.y, n4.z});
    const float dnn = sqrtf(dn.x * dn.x + dn.y * dn.y + dn.z * dn.z);
    double dt = a.dt_d;
    if (dxn > a.dr_opt || dnn > a.dn_opt) dt = fmin(a.dr_opt / dxn, a.dn_opt / dnn);
    const float sx = (float)(dt / m), sn = (float)dt;
    x.x += f.x * sx; x.y += f.y * sx; x.z += f.z * sx;
    const float gx = n4.x + dn.x * sn, gy = n4.y + dn.y * sn, gz = n4.z + dn.z * sn;
    const float s = 1.0f / sqrtf(gx * gx + gy * gy + gz * gz);
    const float4 nnew = make_float4(gx * s, gy * s, gz * s, n4.w);
    if (x.x < a.dlo || x.x > a.dhi || x.y < a.dlo || x.y > a.dhi || x.z < a.dlo || x.z > a.dhi) {   // bounce_back (rare)
        float4 v = a.v[i];
        bounce(x.x, v.x, a.dlo, a.dhi); bounce(x.y, v.y, a.dlo, a.dhi); bounce(x.z, v.z, a.dlo, a.dhi);
        a.v[i] = v;
    }
    a.x_out[i] = x; a.nn_out[i] = nnew;
    if (a.clear) { a.f[i] = make_float4(0, 0, 0, 0); a.t[i] = make_float4(0, 0, 0, 0); }
    else a.t[i] = make_float4(t.x, t.y, t.z, 0);
    if (a.push.world > 1) {
        unsigned msk = a.push.cell_mask[a.push.cellid[i]];
        if (a.push.pmask) msk |= a.push.pmask[i];
        while (msk) {
            const int r = __ffs(msk) - 1; msk &= msk - 1;
            a.push.x[r][i] = x; a.push.nn[r][i] = nnew;
        }
    }
}

// Nose-Hoover friction update of the functor destructors (integrate_nh.h:181-185,240-244), on the device for orbc_run_nh:
// nh[0] = zeta, nh[1] = Q; acc[0] = KE (consumed and reset)
// Decomposed run: ke_all holds one partial kinetic energy per rank (k_share_ke + barrier); they are summed in rank order, so
// every rank arrives at the same zeta.
__global__ void k_nh_zeta_update(float *nh, double *acc, double dt, float kBT, long n, const double *ke_all, int world) {
    if (threadIdx.x || blockIdx.x) return;
    double ke = acc[0];
    if (ke_all) { ke = 0.0; for (int r = 0; r < world; ++r) ke += ke_all[r]; }
    if (!nh[1]) nh[1] = (float)(n * 0.01);
    float zeta = nh[0];
    zeta += 0.5 * dt / nh[1] * (ke - 0.5 * 3.0 * n * kBT);
    nh[0] = zeta;
    acc[0] = 0.0;
    acc[5] = ke;                                                 // the summed kinetic energy, for callers that return it
}
struct KeDst { double *dst[kMaxWorld]; };
__global__ void k_share_ke(const double *acc, int rank, int world, KeDst d) {
    const int r = threadIdx.x;
    if (r < world) d.dst[r][rank] = acc[0];
}
__global__ void k_sum_ke(double *acc, const double *ke_all, int world) {
    if (threadIdx.x || blockIdx.x) return;
    double ke = 0.0; for (int r = 0; r < world; ++r) ke += ke_all[r];
    acc[0] = ke;
}

// ---- constrain_volume.h:26-83 ------------------------------------------------------------------------------------------------
// pass 1: acc[1..3] = centroid sum.  One block, fixed summation order: every rank of a decomposed run holds all the centroids
// and must arrive at the same centre.
__global__ void __launch_bounds__(1024) k_cv_center(const float4 *__restrict__ centroid, int n_cells, double *acc) {
    __shared__ double s[3][1024];
    double x = 0, y = 0, z = 0;
    for (int i = threadIdx.x; i < n_cells; i += 1024) { const float4 c = centroid[i]; x += c.x; y += c.y; z += c.z; }
    s[0][threadIdx.x] = x; s[1][threadIdx.x] = y; s[2][threadIdx.x] = z;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) {
        if (threadIdx.x < d) { s[0][threadIdx.x] += s[0][threadIdx.x + d]; s[1][threadIdx.x] += s[1][threadIdx.x + d]; s[2][threadIdx.x] += s[2][threadIdx.x + d]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { acc[1] = s[0][0]; acc[2] = s[1][0]; acc[3] = s[2][0]; }
}
// pass 2, cells [cb, ce): per cell outward normal (persistent scratch, never cleared — constrain_volume.h:34,55) and volume,
// acc[4] += volume of these cells
__global__ void __launch_bounds__(256) k_cv_normal_volume(const float4 *__restrict__ centroid, int n_cells, int cb, int ce, const int *__restrict__ cs_l, const float4 *__restrict__ nl,
                                                           float4 *__restrict__ cell_normal, double *acc) {
    const int i = cb + blockIdx.x * blockDim.x + threadIdx.x;
    double vol = 0.0;
    if (i < ce) {
        const float cx = (float)acc[1] / n_cells, cy = (float)acc[2] / n_cells, cz = (float)acc[3] / n_cells;
        float4 cn = cell_normal[i];
        const int b = cs_l[i], e = cs_l[i + 1];
        for (int j = b; j < e; ++j) { const float4 q = nl[j]; cn.x += q.x; cn.y += q.y; cn.z += q.z; }
        const float s = 1.0f / sqrtf(cn.x * cn.x + cn.y * cn.y + cn.z * cn.z);
        cn.x *= s; cn.y *= s; cn.z *= s;
        const float4 c = centroid[i];
        const float dx = c.x - cx, dy = c.y - cy, dz = c.z - cz;
        if (cn.x * dx + cn.y * dy + cn.z * dz < 0) { cn.x = -cn.x; cn.y = -cn.y; cn.z = -cn.z; }
        const float height = dx * cn.x + dy * cn.y + dz * cn.z;
        vol = (double)(height * (e - b)) * 3.1415926 * 1.26 / 4.0 / 3.0 * 1e-6;
        cell_normal[i] = cn;
    }
    block_add_double(vol, acc + 4);
}
// decomposed run: this rank's partial volume -> slot `rank` of every rank's vol_all, and the types of the owned protein slots
// below n_cells -> every rank's cv_ptype (pass 3 indexes the protein mass by the CELL index); a barrier follows
struct CvShare { double *vol[kMaxWorld]; int *ptype[kMaxWorld]; };
__global__ void k_cv_share(const double *__restrict__ acc, const float4 *__restrict__ xp, const int *__restrict__ range, int n_cells, int rank, int world, CvShare d) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < world) d.vol[k][rank] = acc[4];
    const int i = range[2] + k;
    if (i < range[3] && i < n_cells) {
        const int t = __float_as_int(xp[i].w);
        for (int r = 0; r < world; ++r) d.ptype[r][i] = t;
    }
}
__global__ void k_sum_partials(double *dst, const double *all, int world) {
    if (threadIdx.x || blockIdx.x) return;
    double s = 0.0; for (int r = 0; r < world; ++r) s += all[r];
    *dst = s;
}
// pass 3: f += strength (V0 - V) / V0 * normal * mass[type[CELL index]] — the reference indexes the mass by the cell index
// (constrain_volume.h:70,73): lipid.type[i] is always 0, prote.type[i] is the type of protein number i.
// slots [range[0], range[1]); the protein type comes from the container (type_src, one GPU) or from cv_ptype (decomposed);
// the volume is acc[4] (one GPU) or the sum of the ranks' partial volumes in rank order (vol_all)
__global__ void k_cv_apply(const int *__restrict__ cellid, const int *__restrict__ range, const float4 *__restrict__ cell_normal, const float4 *__restrict__ type_src, size_t n_type_src,
                           const int *__restrict__ ptype, int is_protein, float target, float strength, const double *acc, const double *vol_all, int world, float4 *__restrict__ f) {
    const size_t j = (size_t)range[0] + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (size_t)range[1]) return;
    const int c = cellid[j];
    double vsum = acc[4];
    if (vol_all) { vsum = 0.0; for (int r = 0; r < world; ++r) vsum += vol_all[r]; }
    const float volume = (float)vsum;
    const float fs = strength * (target - volume) / target;
    float m = c_ff.mass[0];
    if (is_protein) {
        int t = 0;
        if (ptype) t = ptype[c];
        else if ((size_t)c < n_type_src) t = __float_as_int(type_src[c].w);
        m = c_ff.mass[t];
    }
    const float4 cn = cell_normal[c];
    float4 ff = f[j];
    ff.x += fs * cn.x * m; ff.y += fs * cn.y * m; ff.z += fs * cn.z * m;
    f[j] = ff;
}

} // namespace orbc
