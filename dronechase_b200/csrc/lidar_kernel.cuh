// dronechase_b200 -- stand-alone projection-LiDAR kernel (threatsense microbenchmark shape:
// 16 entities per env, several observing wingmen per env, one (C,13,26) sphere per observer).
// One block per env: the env's entity snapshot (position, quaternion, type, alive) is staged in
// shared memory once and reused by every observer x entity projection; the spheres of the env are
// one contiguous slab that the whole block fills with coalesced stores before the winners scatter.
#pragma once
#include "common.cuh"
#include "lidar.cuh"

namespace dc {

constexpr int LIDAR_THREADS = 128;
constexpr int LIDAR_MAX_ENT = 128;
constexpr int LIDAR_MAX_PAIRS = 1024;

__global__ void __launch_bounds__(LIDAR_THREADS)
lidar_kernel(const float* __restrict__ pos, const float* __restrict__ quat, const int32_t* __restrict__ type,
             const uint8_t* __restrict__ alive, const int32_t* __restrict__ obs_slot, int n_ent, int n_obs,
             int flavour, double radius, float* __restrict__ sphere, int32_t* __restrict__ ids) {
    __shared__ float s_pos[LIDAR_MAX_ENT * 3];
    __shared__ float s_quat[LIDAR_MAX_ENT * 4];
    __shared__ int s_type[LIDAR_MAX_ENT];
    __shared__ int s_alive[LIDAR_MAX_ENT];
    __shared__ int s_obs[LIDAR_MAX_ENT];
    __shared__ int s_cell[LIDAR_MAX_PAIRS];
    __shared__ double s_rn[LIDAR_MAX_PAIRS];
    const int env = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < n_ent * 3; i += LIDAR_THREADS) s_pos[i] = pos[(long long)env * n_ent * 3 + i];
    for (int i = tid; i < n_ent * 4; i += LIDAR_THREADS) s_quat[i] = quat[(long long)env * n_ent * 4 + i];
    for (int i = tid; i < n_ent; i += LIDAR_THREADS) { s_type[i] = type[i]; s_alive[i] = alive[(long long)env * n_ent + i]; }
    for (int i = tid; i < n_obs; i += LIDAR_THREADS) s_obs[i] = obs_slot[i];
    __syncthreads();

    const int ch = flavour == 0 ? 3 : 2;
    const int per = ch * N_CELLS;
    const int chunk = LIDAR_MAX_PAIRS / n_ent;            // observers handled per pass
    for (int o0 = 0; o0 < n_obs; o0 += chunk) {
        const int no = min(chunk, n_obs - o0);
        // empty spheres first: the stores are the kernel's HBM traffic (4 KB per sphere) and drain while the projections
        // below are computed.  ch * 338 floats per sphere is even -> 8-byte stores are always aligned; 16-byte stores when
        // the slab of this pass starts on a 16-byte boundary and holds a multiple of four floats.
        float* slab = sphere + ((long long)env * n_obs + o0) * per;
        const int n_fl = no * per;
        if ((reinterpret_cast<uintptr_t>(slab) & 15) == 0 && (n_fl & 3) == 0) {
            float4* out4 = reinterpret_cast<float4*>(slab);
            for (int i = tid; i < n_fl / 4; i += LIDAR_THREADS) out4[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        } else {
            float2* out2 = reinterpret_cast<float2*>(slab);
            for (int i = tid; i < n_fl / 2; i += LIDAR_THREADS) out2[i] = make_float2(1.f, 1.f);
        }
        if (ids) {
            int32_t* idp = ids + ((long long)env * n_obs + o0) * N_CELLS;
            for (int i = tid; i < no * N_CELLS; i += LIDAR_THREADS) idp[i] = -1;
        }
        for (int pr = tid; pr < no * n_ent; pr += LIDAR_THREADS) {
            const int o = pr / n_ent, k = pr - o * n_ent, ob = s_obs[o0 + o];
            int cell = -1; double rn = 1.0;
            if (k != ob && s_alive[k] && s_alive[ob]) {
                if (flavour == 0) {
                    // fused: only (cell, r_n) are kept -> float32 angles with the float64 path near a cell border,
                    // the same cell the reference's float64 arithmetic picks (lidar.cuh, lidar_cell_fused)
                    lidar_cell_fused(radius, s_pos[3 * ob], s_pos[3 * ob + 1], s_pos[3 * ob + 2],
                                     s_quat[4 * ob], s_quat[4 * ob + 1], s_quat[4 * ob + 2], s_quat[4 * ob + 3],
                                     s_pos[3 * k], s_pos[3 * k + 1], s_pos[3 * k + 2], &cell, &rn);
                } else {
                    LidarHit h = lidar_project_one(flavour, radius, s_pos[3 * ob], s_pos[3 * ob + 1], s_pos[3 * ob + 2],
                                                   s_quat[4 * ob], s_quat[4 * ob + 1], s_quat[4 * ob + 2], s_quat[4 * ob + 3],
                                                   s_pos[3 * k], s_pos[3 * k + 1], s_pos[3 * k + 2]);
                    cell = h.cell; rn = h.rn;
                }
            }
            s_cell[pr] = cell; s_rn[pr] = rn;
        }
        __syncthreads();
        for (int pr = tid; pr < no * n_ent; pr += LIDAR_THREADS) {
            const int o = pr / n_ent, k = pr - o * n_ent;
            if (!lidar_wins(flavour, k, n_ent, s_cell + o * n_ent, s_rn + o * n_ent)) continue;
            float* sph = sphere + ((long long)env * n_obs + o0 + o) * per;
            const int c = s_cell[pr];
            sph[c] = (float)s_rn[pr];
            sph[N_CELLS + c] = (float)((double)s_type[k] / 5.0);
            if (ch == 3) sph[2 * N_CELLS + c] = 0.1f;
            if (ids) ids[((long long)env * n_obs + o0 + o) * N_CELLS + c] = k;
        }
        __syncthreads();
    }
}

}  // namespace dc

namespace dc {

// ------------------------------------------------------------------------------------------------
// Ray-cast variant (opt-in sensor model; the reference only has the projection above).  One ray per
// cell through the cell centre (theta_c, phi_c of LidarMath.radian_from_index, lidar_math.py:103-105)
// is tested against every entity's bounding sphere staged in shared memory (the rayTestBatch
// replacement); the nearest hit wins.  Every entity additionally claims the cell that contains its
// centre at its centre distance -- so a cell never reads farther than the projection would, occlusion
// by a nearer body works, and with radii -> 0 the result reduces to the reference's projection sphere.
// ------------------------------------------------------------------------------------------------
constexpr int RAY_THREADS = 128;

__global__ void __launch_bounds__(RAY_THREADS)
raycast_kernel(const float* __restrict__ pos, const float* __restrict__ quat, const float* __restrict__ ent_radius,
               const int32_t* __restrict__ type, const uint8_t* __restrict__ alive, const int32_t* __restrict__ obs_slot,
               int n_ent, int n_obs, double max_range, float* __restrict__ sphere, int32_t* __restrict__ ids) {
    __shared__ float s_pos[LIDAR_MAX_ENT * 3];
    __shared__ float s_quat[LIDAR_MAX_ENT * 4];
    __shared__ float s_rad[LIDAR_MAX_ENT];
    __shared__ int s_type[LIDAR_MAX_ENT];
    __shared__ int s_alive[LIDAR_MAX_ENT];
    __shared__ int s_cell[LIDAR_MAX_ENT];      // projection cell of every entity for the current observer
    __shared__ float s_dist[LIDAR_MAX_ENT];    // centre distance (normalised)
    const int env = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < n_ent * 3; i += RAY_THREADS) s_pos[i] = pos[(long long)env * n_ent * 3 + i];
    for (int i = tid; i < n_ent * 4; i += RAY_THREADS) s_quat[i] = quat[(long long)env * n_ent * 4 + i];
    for (int i = tid; i < n_ent; i += RAY_THREADS) {
        s_type[i] = type[i]; s_alive[i] = alive[(long long)env * n_ent + i]; s_rad[i] = ent_radius[i];
    }
    __syncthreads();
    const float PI_F = 3.14159265358979f;
    for (int o = 0; o < n_obs; ++o) {
        const int ob = obs_slot[o];
        // projection of the centres (float64, same arithmetic as the reference-pinned projection)
        for (int k = tid; k < n_ent; k += RAY_THREADS) {
            int cell = -1; float dn = 1.0f;
            if (k != ob && s_alive[k] && s_alive[ob]) {
                LidarHit h = lidar_project_one(0, max_range, s_pos[3 * ob], s_pos[3 * ob + 1], s_pos[3 * ob + 2], s_quat[4 * ob],
                                               s_quat[4 * ob + 1], s_quat[4 * ob + 2], s_quat[4 * ob + 3], s_pos[3 * k],
                                               s_pos[3 * k + 1], s_pos[3 * k + 2]);
                cell = h.cell; dn = (float)h.rn;
            }
            s_cell[k] = cell; s_dist[k] = dn;
        }
        __syncthreads();
        const float qx = s_quat[4 * ob], qy = s_quat[4 * ob + 1], qz = s_quat[4 * ob + 2], qw = s_quat[4 * ob + 3];
        const float r00 = 1 - 2 * (qy * qy + qz * qz), r01 = 2 * (qx * qy - qw * qz), r02 = 2 * (qx * qz + qw * qy);
        const float r10 = 2 * (qx * qy + qw * qz), r11 = 1 - 2 * (qx * qx + qz * qz), r12 = 2 * (qy * qz - qw * qx);
        const float r20 = 2 * (qx * qz - qw * qy), r21 = 2 * (qy * qz + qw * qx), r22 = 1 - 2 * (qx * qx + qy * qy);
        float* sph = sphere + ((long long)env * n_obs + o) * 3 * N_CELLS;
        int32_t* idp = ids ? ids + ((long long)env * n_obs + o) * N_CELLS : nullptr;
        for (int c = tid; c < N_CELLS; c += RAY_THREADS) {
            const int ti = c / N_PHI, pj = c - ti * N_PHI;
            const float th = (ti + 0.5f) / N_THETA * PI_F, ph = -PI_F + (pj + 0.5f) / N_PHI * 2.0f * PI_F;
            float st, ct, sp, cp;
            sincosf(th, &st, &ct); sincosf(ph, &sp, &cp);
            const float bx = st * cp, by = st * sp, bz = ct;                 // body-frame ray
            const float dx = r00 * bx + r01 * by + r02 * bz, dy = r10 * bx + r11 * by + r12 * bz,
                        dz = r20 * bx + r21 * by + r22 * bz;                  // world-frame ray
            float best = 1.0f; int best_id = -1;
            if (s_alive[ob]) {
                for (int k = 0; k < n_ent; ++k) {
                    if (k == ob || !s_alive[k]) continue;
                    const float ox = s_pos[3 * k] - s_pos[3 * ob], oy = s_pos[3 * k + 1] - s_pos[3 * ob + 1],
                                oz = s_pos[3 * k + 2] - s_pos[3 * ob + 2];
                    const float tca = ox * dx + oy * dy + oz * dz;
                    const float d2 = ox * ox + oy * oy + oz * oz - tca * tca;
                    const float r2 = s_rad[k] * s_rad[k];
                    float dn = 2.0f;
                    if (tca > 0.0f && d2 <= r2) dn = fminf(fmaxf((tca - sqrtf(r2 - d2)) / (float)max_range, 0.0f), 1.0f);
                    if (s_cell[k] == c) dn = fminf(dn, s_dist[k]);           // the centre always marks its own cell
                    if (dn < best) { best = dn; best_id = k; }
                }
            }
            sph[c] = best;
            sph[N_CELLS + c] = best_id >= 0 ? (float)((double)s_type[best_id] / 5.0) : 1.0f;
            sph[2 * N_CELLS + c] = best_id >= 0 ? 0.1f : 1.0f;
            if (idp) idp[c] = best_id;
        }
        __syncthreads();
    }
}

// Dense (C,13,26) spheres from hit lists (dc_buffers.lidar_hits rows: per entity slot (cell, float bits of r_n), cell = -1
// for none) -- the device-side twin of dc_host_scatter_sphere, used to rebuild observations of a rollout that was stored
// sparsely (56 B instead of 4 KB per env step for D = 7).  One block per output row; row_index (optional) gathers rows
// of a [T * E, D, 2] rollout buffer into a minibatch.
constexpr int SCATTER_THREADS = 128;
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_hits_kernel(const int2* __restrict__ hits, const long long* __restrict__ row_index, int n_drones, int n_lw,
                    int channels, float* __restrict__ dense) {
    const long long row = blockIdx.x;
    const long long src = row_index ? row_index[row] : row;
    const int per = channels * N_CELLS;
    float2* out2 = reinterpret_cast<float2*>(dense + row * per);       // per is even: 8-byte stores stay aligned
    for (int i = threadIdx.x; i < per / 2; i += SCATTER_THREADS) out2[i] = make_float2(1.f, 1.f);
    __syncthreads();
    float* sph = dense + row * per;
    for (int d = threadIdx.x; d < n_drones; d += SCATTER_THREADS) {
        const int2 h = hits[src * n_drones + d];
        if (h.x < 0 || h.x >= N_CELLS) continue;
        sph[h.x] = __int_as_float(h.y);
        sph[N_CELLS + h.x] = d < n_lw ? 0.6f : 0.2f;                  // EntityType value / 5
        if (channels == 3) sph[2 * N_CELLS + h.x] = 0.1f;               // normalised age 1/10
    }
}

// Same for the level5 stacked observation: dense[row] (6,3,13,26) = six empty spheres + the hit list of row
// (row_index ? row_index[row] : row) of `hits` ([rows, cap, 2], cap = 5 * n_drones + 1, the level5 layout of
// dc_buffers.lidar_hits / student_hits: code = sphere * 338 + cell | wingman << 11 | age << 12, terminated by -1).
// Device-side twin of dc_host_scatter_stack.
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_stack_kernel(const int2* __restrict__ hits, const long long* __restrict__ row_index, int cap,
                     float* __restrict__ dense) {
    const long long row = blockIdx.x;
    const long long src = row_index ? row_index[row] : row;
    constexpr int per = 6 * 3 * N_CELLS;                               // 6084 floats: 16-byte stores stay aligned
    float4* out4 = reinterpret_cast<float4*>(dense + row * per);
    for (int i = threadIdx.x; i < per / 4; i += SCATTER_THREADS) out4[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    __syncthreads();
    float* st = dense + row * per;
    const int2* list = hits + src * cap;
    // the list is -1 terminated: a lane past the terminator sees codes of an older, longer list -> find the end first
    __shared__ int s_end;
    if (threadIdx.x == 0) s_end = cap;
    __syncthreads();
    for (int i = threadIdx.x; i < cap; i += SCATTER_THREADS)
        if (list[i].x < 0) atomicMin(&s_end, i);
    __syncthreads();
    const int n = s_end;
    for (int i = threadIdx.x; i < n; i += SCATTER_THREADS) {
        const int2 h = list[i];
        const int code = h.x & 2047, sp = code / N_CELLS, c = code - sp * N_CELLS, age = (h.x >> 12) & 15;
        float* o = st + sp * 3 * N_CELLS + c;
        o[0] = __int_as_float(h.y);
        o[N_CELLS] = (h.x >> 11 & 1) ? 0.6f : 0.2f;                     // EntityType value / 5
        o[2 * N_CELLS] = age == 0 ? 0.1f : (float)((double)age / 10.0);  // own sphere 0.1, neighbour snapshot age / RING
    }
}

// Mirror of the incrementally maintained sphere in a SECOND dense array -- in practice the numpy observation array of the
// SB3 adapter, page-locked host memory mapped into the device address space (dc_host_register): the kernel's stores go
// over PCIe as posted writes, a handful of 4-byte words per env, so neither the 4 KB sphere nor its hit list has to be
// copied and no host core has to scatter anything (dc_host_scatter_sphere is the CPU twin, same arithmetic).
//   shown [E,D,2]: the hits `dense` currently shows (updated to `hits` by this kernel); hits [E,D,2]: dc_buffers.lidar_hits.
// One thread per env, slots in order: the cells of `shown` that are no longer held go back to 1.0 first (two slots of an
// env may name the same cell), then the new hits are written; a slot that keeps its cell only refreshes the distance.
constexpr int MIRROR_THREADS = 128;
__global__ void __launch_bounds__(MIRROR_THREADS)
mirror_hits_kernel(int2* __restrict__ shown, const int2* __restrict__ hits, int n_envs, int n_drones, int n_lw, int channels,
                   float* __restrict__ dense) {
    const int e = blockIdx.x * MIRROR_THREADS + threadIdx.x;
    if (e >= n_envs) return;
    int2* p = shown + (long long)e * n_drones;
    const int2* h = hits + (long long)e * n_drones;
    float* sph = dense + (long long)e * channels * N_CELLS;
    for (int d = 0; d < n_drones; ++d) {
        const int c = p[d].x;
        if (c < 0 || c >= N_CELLS || c == h[d].x) continue;
        sph[c] = 1.0f; sph[N_CELLS + c] = 1.0f;
        if (channels == 3) sph[2 * N_CELLS + c] = 1.0f;
    }
    for (int d = 0; d < n_drones; ++d) {
        const int2 hv = h[d];
        const int2 pv = p[d];
        if (hv.x == pv.x && hv.y == pv.y) continue;                       // nothing moved (also: both empty)
        p[d] = hv;
        if (hv.x < 0 || hv.x >= N_CELLS) continue;
        sph[hv.x] = __int_as_float(hv.y);
        if (hv.x == pv.x) continue;
        sph[N_CELLS + hv.x] = d < n_lw ? 0.6f : 0.2f;                     // EntityType value / 5
        if (channels == 3) sph[2 * N_CELLS + hv.x] = 0.1f;                 // normalised age 1/10
    }
}

// The same update as mirror_hits_kernel, but as a LIST the host applies: (flat float index into `dense` [E, C, 13, 26],
// float bits) pairs appended to out[1..], their number in out[0].x.  Posted 4-byte writes into mapped host memory run at
// ~0.3 G/s (0.42 ms per 65,536 envs); a list of the same words crosses PCIe on the copy engines in tens of microseconds
// and a few host threads store it (dc_host_apply_pairs).  An un-write whose cell is entered by another slot in the same
// step is dropped -- the entering slot's three words overwrite it -- so the pairs of one call never repeat an address and
// may be applied in any order by any number of threads.  One thread per env; a warp reserves its range with one atomic
// and every env's pairs are contiguous.  The caller sizes `out` for the worst case (6 words per drone slot + the header).
__global__ void __launch_bounds__(MIRROR_THREADS)
diff_hits_kernel(int2* __restrict__ shown, const int2* __restrict__ hits, int n_envs, int n_drones, int n_lw, int channels,
                 int2* __restrict__ out) {
    const int e = blockIdx.x * MIRROR_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int n = 0;
    int2* p = nullptr; const int2* h = nullptr;
    auto entered = [&](int c, int except) {               // does a slot other than `except` hold cell c in the new hits?
        for (int k = 0; k < n_drones; ++k) if (k != except && h[k].x == c) return true;
        return false;
    };
    if (e < n_envs) {
        p = shown + (long long)e * n_drones;
        h = hits + (long long)e * n_drones;
        for (int d = 0; d < n_drones; ++d) {              // pass 1: count
            const int2 pv = p[d], hv = h[d];
            if (pv.x >= 0 && pv.x < N_CELLS && pv.x != hv.x && !entered(pv.x, d)) n += channels;
            if (hv.x == pv.x && hv.y == pv.y) continue;
            if (hv.x < 0 || hv.x >= N_CELLS) continue;
            n += (hv.x == pv.x) ? 1 : channels;
        }
    }
    // warp-level exclusive scan of n, one atomic per warp
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(&out[0].x, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - n;
    if (e >= n_envs || n == 0) {
        if (e < n_envs) for (int d = 0; d < n_drones; ++d) p[d] = h[d];
        return;
    }
    int2* o = out + 1 + base;
    const int e_base = e * channels * N_CELLS;
    const int one = __float_as_int(1.0f);
    for (int d = 0; d < n_drones; ++d) {                  // pass 2: un-writes (against the old `shown`)
        const int2 pv = p[d], hv = h[d];
        if (pv.x >= 0 && pv.x < N_CELLS && pv.x != hv.x && !entered(pv.x, d)) {
            *o++ = make_int2(e_base + pv.x, one); *o++ = make_int2(e_base + N_CELLS + pv.x, one);
            if (channels == 3) *o++ = make_int2(e_base + 2 * N_CELLS + pv.x, one);
        }
    }
    for (int d = 0; d < n_drones; ++d) {                  // writes, and `shown` := `hits`
        const int2 pv = p[d], hv = h[d];
        if (hv.x == pv.x && hv.y == pv.y) continue;
        p[d] = hv;
        if (hv.x < 0 || hv.x >= N_CELLS) continue;
        *o++ = make_int2(e_base + hv.x, hv.y);
        if (hv.x == pv.x) continue;
        *o++ = make_int2(e_base + N_CELLS + hv.x, __float_as_int(d < n_lw ? 0.6f : 0.2f));      // EntityType value / 5
        if (channels == 3) *o++ = make_int2(e_base + 2 * N_CELLS + hv.x, __float_as_int(0.1f));  // normalised age 1/10
    }
}

}  // namespace dc
