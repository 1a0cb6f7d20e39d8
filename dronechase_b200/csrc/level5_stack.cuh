// dronechase_b200 -- threatsense level5: the stacked-sphere observation (third launch of a level5 step).
//
//   FusedLIDAR.read_data            fused_lidar.py:223-251   own sphere + neighbour spheres, padded to 6, shuffled
//   bootstrap / _build_valid_spheres :73-109                 n ~ choice(1..4) wingman publishers, one random age 1..9 each
//   SnapshotBuffer.get_random_*      lidar_buffer.py:104-145
//   LidarMath.transform_features     lidar_math.py:186-260    denormalise, neighbour frame -> world -> observer frame
//   add_features(invert=True)        lidar_math.py:262-311    farther wins, anything beats an empty cell
//   _pad_sphere_stack / randomize_stack  fused_lidar.py:253-269,293-326
//
// One warp per env.  env_kernel has already written this step's ring entry of every armed wingman (its float32 pose
// and its kept features) and the per-env stack mode.  What the agent's ring provably holds (oracle/level5_oracle.py,
// pinned against the reference's own classes): the snapshot created at ring step s carries wingman P's pose of step
// s + 1 and -- for a neighbour -- the features P broadcast at step s (none at s = 0), for the observer itself the
// features of step s + 1.  The fusion draws are the FUSE Philox stream, index 16 * obs_call + {0: n, 1..4: sample,
// 5..8: ages, 9..13: shuffle}, sub = the agent's slot.
//
// Work layout: the kept features of all (at most five) source snapshots are first compacted into one per-warp item
// list, so the float64 re-framing (two sincos, two rotations, acos, atan2) runs with one feature per lane instead of
// one pass per snapshot with a handful of live lanes (r1m profile: 3.4 of 32 lanes, 395 us -> see profiles/).
//
// The (6,3,13,26) observation lives in the caller's tensor across steps and is maintained incrementally like the
// level4 sphere: the cells marked by the previous step (stack_prev) go back to 1.0, then the new hits are written --
// 24 KB per env would otherwise be rewritten for a few dozen marked cells.
#pragma once
#include "stage03.cuh"

namespace dc {

constexpr int STACK_WARPS = 4;
constexpr int STACK_MAX_D = 64;                       // level5 envs hold at most 64 drones (dc_create checks)
constexpr int STACK_MAX_SRC = 5;                      // own + n_neighbors_max - 1 drawn snapshots
constexpr int STACK_MAX_ITEMS = STACK_MAX_SRC * STACK_MAX_D;

__device__ __forceinline__ int nib_get(uint32_t v, int i) { return (v >> (4 * i)) & 15; }
__device__ __forceinline__ uint32_t nib_set(uint32_t v, int i, int x) { return (v & ~(15u << (4 * i))) | ((uint32_t)x << (4 * i)); }

// VARIANT 1 (STACK_STUDENT) = the stack behind info["student_observation"] of the base env
// (level5_envrionment.py:291-292,342-346): the env's SECOND compute_observation call of the step.  Same ring (its
// updates are idempotent), the draws of obs_call + 1, and every wingman is a candidate (the dead ones all re-opened
// their buffer during the first call).  The host passes the student tensors as A.obs_lidar / A.obs_mask /
// A.p.stack_prev; the count of marked cells lives in A.p.mo_prev_n.
// VARIANT 2 (STACK_MULTI) = Level5DumbMultiObs.compute_info (level5_dumb_multiobs.py:112-150): one stack per (env,
// wingman) -- every ARMED wingman is an observer, the FUSE draws are keyed by its slot, the candidates are the armed
// wingmen; a disarmed observer's stack is emptied.  Tensors are [E][n_lw][...], one warp per (env, observer).
enum { STACK_MAIN = 0, STACK_STUDENT = 1, STACK_MULTI = 2 };
template <typename R, int VARIANT>
__global__ void __launch_bounds__(STACK_WARPS * 32) stack_kernel(const StepArgs<R> A) {
    constexpr bool STUDENT = VARIANT == STACK_STUDENT, MULTI = VARIANT == STACK_MULTI;
    __shared__ int s_item[STACK_WARPS][STACK_MAX_ITEMS];       // src << 8 | entity slot
    __shared__ int s_cell[STACK_WARPS][STACK_MAX_ITEMS];
    __shared__ double s_rn[STACK_WARPS][STACK_MAX_ITEMS];
    __shared__ int s_begin[STACK_WARPS][STACK_MAX_SRC + 1];    // first item of every source
    __shared__ int s_desc[STACK_WARPS][6];                     // n_items, agent, cur, srcP, srcAge of the warp's env
    __shared__ __align__(16) float s_pose[STACK_WARPS][8];     // the observer's float32 pose
    static_assert(STACK_WARPS == 4, "the shared re-framing pass indexes four item lists");
    const TaskParams& T = A.t;
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = T.D, L = T.n_lw;
    const int n_obs = MULTI ? T.n_envs * L : T.n_envs;             // observers: (env, wingman) pairs or envs
    const int item_raw = blockIdx.x * STACK_WARPS + wi;
    const int item = item_raw < n_obs ? item_raw : n_obs - 1;     // surplus warps of the last block idle behind `build`
    const int env = MULTI ? item / L : item;
    int32_t* w5 = A.p.env5 + (long long)env * ENV5_WORDS;
    const int my_flag = lane < L ? A.p.flagw[(long long)env * D + lane] : 0;
    int* prev_n_word = (MULTI || STUDENT) ? A.p.mo_prev_n + item : w5 + W5_PREV_N;
    const int mode = item_raw < n_obs ? (w5[W5_STACK_MODE] & 255) : STACK_KEEP, cand_mask = w5[W5_STACK_MODE] >> 8;
    // The re-framing below is shared by the four warps of the block (their few items together fill a warp), so no
    // warp leaves early: `build` says whether this warp's env gets a new stack at all.
    bool build = mode != STACK_KEEP;
    float* obs = A.obs_lidar + (long long)item * N_STACK * 3 * N_CELLS;
    // hit list of the stacked observation: (code, float bits of r_n) per marked cell, code = sphere * 338 + cell |
    // wingman << 11 | age << 12 (age 0 = the observer's own sphere), terminated by code = -1 (dc_buffers.lidar_hits)
    const int cap = STACK_MAX_SRC * D + 1;
    int2* prev = A.p.stack_prev + (long long)item * cap;
    const int prev_n = build ? *prev_n_word : 0;
    for (int i = lane; i < prev_n; i += 32) {
        const int code = prev[i].x & 2047, sp = code / N_CELLS, c = code - sp * N_CELLS;
        float* o = obs + sp * 3 * N_CELLS + c;
        o[0] = 1.0f; o[N_CELLS] = 1.0f; o[2 * N_CELLS] = 1.0f;
    }
    uint8_t* mask = A.obs_mask + (long long)item * N_STACK;
    const int ag = MULTI ? item - env * L : w5[W5_AGENT];
    // candidates: wingmen still publishing (never disarmed in this episode), in slot order; 4-bit fields.  The flag
    // words were requested at the top of the kernel; the agent-centred variants vote after the Philox block, which
    // hides the rest of that load's latency
    unsigned armed_bits = 0;
    if (MULTI) armed_bits = __ballot_sync(0xffffffffu, (my_flag & F_ARMED) != 0);
    // reset observation: the ring was wiped by the step-0 broadcast; (multi) a disarmed wingman observes nothing
    if (mode == STACK_EMPTY || (MULTI && build && !(armed_bits >> ag & 1))) {
        if (lane < N_STACK) mask[lane] = 0;
        if (lane == 0) { *prev_n_word = 0; prev[0] = make_int2(-1, 0); }
        build = false;
    }
    const int cur = A.p.env[(long long)env * ENV_WORDS + W_STEP];
    const uint32_t call = (uint32_t)(w5[W5_OBS_CALL] - (T.l5_base ? 3 : 1) + (STUDENT ? 1 : 0));   // base env: the first (student: second) of the step's three calls
    const double my_u = lane < 14 ? philox_uniform(T.k0, T.k1, T.env_offset + (uint32_t)env, STREAM_FUSE, 16u * call + (uint32_t)lane, (uint32_t)ag) : 0.0;
    auto u = [&](int i) { return __shfl_sync(0xffffffffu, my_u, i); };
    if (!MULTI) armed_bits = __ballot_sync(0xffffffffu, (my_flag & F_ARMED) != 0);
    uint32_t cands = 0; int m = 0;
    for (int P = 0; P < L; ++P) if (STUDENT || ((T.l5_base ? (unsigned)cand_mask : armed_bits) >> P & 1)) cands = nib_set(cands, m++, P);
    const int n = 1 + (int)(u(0) * 4.0);
    const int k = n < m ? n : m;
    for (int i = 0; i < k; ++i) {             // random.sample: partial Fisher-Yates
        const int j = i + (int)(u(1 + i) * (double)(m - i));
        const int a = nib_get(cands, i), b = nib_get(cands, j);
        cands = nib_set(nib_set(cands, i, b), j, a);
    }
    // sphere sources: 0 = own, then the drawn snapshots that exist (age <= steps since the reset); 4-bit fields
    uint32_t srcP = (uint32_t)ag, srcAge = 0; int n_src = 1;
    for (int i = 0; i < k; ++i) {
        const int a = 1 + (int)(u(5 + i) * 9.0);
        if (cur - a < 0) continue;
        if (!(armed_bits >> nib_get(cands, i) & 1)) continue;     // (base env) re-registered after its death: no pose, no sphere
        srcP = nib_set(srcP, n_src, nib_get(cands, i)); srcAge = nib_set(srcAge, n_src, a); ++n_src;
    }
    uint32_t order = 0x543210u;
    for (int kk = 0, i = N_STACK - 1; i > 0; --i, ++kk) {   // random.shuffle
        const int j = (int)(u(9 + kk) * (double)(i + 1));
        const int a = nib_get(order, i), b = nib_get(order, j);
        order = nib_set(nib_set(order, i, b), j, a);
    }
    uint32_t dst_of = 0;
    for (int dst = 0; dst < N_STACK; ++dst) dst_of = nib_set(dst_of, nib_get(order, dst), dst);
    if (build && lane < N_STACK) mask[lane] = nib_get(order, lane) < n_src ? 1 : 0;
    if (!build) n_src = 0;

    // ring entry holding the features of source `src`: own -> this step; the observer drawn as its own neighbour ->
    // step s + 1; another wingman -> step s (nothing at s = 0: the reset observation broadcast no features)
    auto entry_of = [L](int env_, int ag_, int cur_, uint32_t srcP_, uint32_t srcAge_, int src) -> long long {
        const int P = nib_get(srcP_, src), s = cur_ - nib_get(srcAge_, src);
        if (src == 0) return ((long long)env_ * L + ag_) * RING + cur_ % RING;
        if (P == ag_) return ((long long)env_ * L + ag_) * RING + (s + 1) % RING;
        return s >= 1 ? ((long long)env_ * L + P) * RING + s % RING : -1;
    };
    auto feat_entry = [&](int src) -> long long { return entry_of(env, ag, cur, srcP, srcAge, src); };
    // ---- item list: every kept feature of every source, grouped by source ----
    // (all ring_meta words of all sources are fetched before the first ballot: one round trip instead of five, and the
    // observer's pose is requested now, long before the re-framing needs it)
    const float* own_pose = A.p.ring_pose + (((long long)env * L + ag) * RING + cur % RING) * 8;
    const float4 own_a = reinterpret_cast<const float4*>(own_pose)[0], own_b = reinterpret_cast<const float4*>(own_pose)[1];
    int metas[STACK_MAX_SRC][STACK_MAX_D / 32];
#pragma unroll
    for (int src = 0; src < STACK_MAX_SRC; ++src) {
        const long long ef = src < n_src ? feat_entry(src) : -1;
#pragma unroll
        for (int c = 0; c < STACK_MAX_D / 32; ++c) {
            const int d = 32 * c + lane;
            metas[src][c] = (ef >= 0 && d < D) ? A.p.ring_meta[ef * D + d] : -1;
        }
    }
    int n_items = 0;
#pragma unroll
    for (int src = 0; src < STACK_MAX_SRC; ++src) {
        if (src >= n_src) break;
        if (lane == 0) s_begin[wi][src] = n_items;
#pragma unroll
        for (int c = 0; c < STACK_MAX_D / 32; ++c) {
            if (32 * c >= D) break;
            const int d = 32 * c + lane;
            const bool valid = metas[src][c] >= 0 && (src == 0 || d != ag);
            const unsigned bal = __ballot_sync(0xffffffffu, valid);
            if (valid) s_item[wi][n_items + __popc(bal & ((1u << lane) - 1))] = (src << 8) | d;
            n_items += __popc(bal);
        }
    }
    if (lane == 0) {
        s_begin[wi][n_src] = n_items;
        s_desc[wi][0] = n_items; s_desc[wi][1] = ag; s_desc[wi][2] = cur; s_desc[wi][3] = (int)srcP; s_desc[wi][4] = (int)srcAge;
        s_desc[wi][5] = env;
        reinterpret_cast<float4*>(s_pose[wi])[0] = own_a; reinterpret_cast<float4*>(s_pose[wi])[1] = own_b;
    }
    __syncthreads();
    const double radius = 2 * T.dome;
    // ---- one feature per thread, over the items of all four envs of the block: re-frame into the observer's frame ----
    // (an env keeps a handful of features: per warp this float64 section ran with 4 of 32 lanes busy, r1u profile)
    {
        const int n0 = s_desc[0][0], n1 = n0 + s_desc[1][0], n2 = n1 + s_desc[2][0], n3 = n2 + s_desc[3][0];
        for (int f = threadIdx.x; f < n3; f += STACK_WARPS * 32) {
            const int w = f < n0 ? 0 : f < n1 ? 1 : f < n2 ? 2 : 3;
            const int i = f - (w == 0 ? 0 : w == 1 ? n0 : w == 2 ? n1 : n2);
            const int ag_ = s_desc[w][1], cur_ = s_desc[w][2], env_ = s_desc[w][5];
            const uint32_t srcP_ = (uint32_t)s_desc[w][3], srcAge_ = (uint32_t)s_desc[w][4];
            const int it = s_item[w][i], src = it >> 8, d = it & 255;
            const long long ef = entry_of(env_, ag_, cur_, srcP_, srcAge_, src);
            const double* ft = A.p.ring_feat + (ef * D + d) * 3;
            int cell; double rn;
            if (src == 0) { cell = A.p.ring_meta[ef * D + d] & 0xffff; rn = ft[0]; }
            else {
                const int P = nib_get(srcP_, src), sn = cur_ - nib_get(srcAge_, src);
                const float* np_ = A.p.ring_pose + (((long long)env_ * L + P) * RING + (sn + 1) % RING) * 8;   // pose of step s + 1
                const double x = np_[3], y = np_[4], z = np_[5], w_ = np_[6];
                const double Rr = ft[0] * radius;
                double st, ct, sp, cp;
                sincos(ft[1], &st, &ct); sincos(ft[2], &sp, &cp);
                const double cx = Rr * st * cp, cy = Rr * st * sp, cz = Rr * ct;
                const double gx = ((1 - 2 * (y * y + z * z)) * cx + 2 * (x * y - w_ * z) * cy + 2 * (x * z + w_ * y) * cz) + (double)np_[0];
                const double gy = (2 * (x * y + w_ * z) * cx + (1 - 2 * (x * x + z * z)) * cy + 2 * (y * z - w_ * x) * cz) + (double)np_[1];
                const double gz = (2 * (x * z - w_ * y) * cx + 2 * (y * z + w_ * x) * cy + (1 - 2 * (x * x + y * y)) * cz) + (double)np_[2];
                const float* op = s_pose[w];
                // only (cell, r_n) are kept of the re-framed feature: float32 angles decide the cell unless one lies within
                // 2e-3 of a cell border (then float64, as the reference computes it), r_n stays float64
                lidar_cell_fused(radius, (double)op[0], (double)op[1], (double)op[2], (double)op[3], (double)op[4], (double)op[5],
                                 (double)op[6], gx, gy, gz, &cell, &rn);
            }
            s_cell[w][i] = cell; s_rn[w][i] = rn;
        }
    }
    __syncthreads();
    if (!build) return;
    // ---- winners (sequential add_features rule inside each source) and the incremental write ----
    int n_new = 0;
    for (int i0 = 0; i0 < n_items; i0 += 32) {
        const int i = i0 + lane;
        bool win = false; int c = -1, src = 0, d = 0;
        if (i < n_items) {
            const int it = s_item[wi][i];
            src = it >> 8; d = it & 255; c = s_cell[wi][i];
            if (src == 0) win = true;          // kept features are one per cell already
            else {
                float curv = 1.0f; int wj = -1;
                for (int j = s_begin[wi][src], je = s_begin[wi][src + 1]; j < je; ++j) {
                    if (s_cell[wi][j] != c) continue;
                    const double rj = s_rn[wi][j];
                    const bool take = curv < 1.0f ? (rj > (double)curv) : true;
                    if (take) { curv = (float)rj; wj = j; }
                }
                win = wj == i;
            }
        }
        const int dst = nib_get(dst_of, src);
        if (win) {
            float* o = obs + dst * 3 * N_CELLS;
            o[c] = (float)s_rn[wi][i];
            o[N_CELLS + c] = d < L ? 0.6f : 0.2f;                                        // EntityType value / 5
            o[2 * N_CELLS + c] = src == 0 ? 0.1f : (float)((double)nib_get(srcAge, src) / RING);   // normalised age
        }
        const unsigned bal = __ballot_sync(0xffffffffu, win);
        if (win) prev[n_new + __popc(bal & ((1u << lane) - 1))] =
            make_int2((dst * N_CELLS + c) | (d < L ? 1 << 11 : 0) | (nib_get(srcAge, src) << 12), __float_as_int((float)s_rn[wi][i]));
        n_new += __popc(bal);
    }
    if (lane == 0) { *prev_n_word = n_new; prev[n_new] = make_int2(-1, 0); }
}

}  // namespace dc
