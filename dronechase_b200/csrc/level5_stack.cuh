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
// The (6,3,13,26) observation lives in the caller's tensor across steps and is maintained incrementally like the
// level4 sphere: the cells marked by the previous step (stack_prev) go back to 1.0, then the new hits are written --
// 24 KB per env would otherwise be rewritten for a few dozen marked cells.
#pragma once
#include "stage03.cuh"

namespace dc {

constexpr int STACK_WARPS = 4;
constexpr int STACK_MAX_D = 256;

template <typename R>
__global__ void __launch_bounds__(STACK_WARPS * 32) stack_kernel(const StepArgs<R> A) {
    __shared__ int s_cell[STACK_WARPS][STACK_MAX_D];
    __shared__ double s_rn[STACK_WARPS][STACK_MAX_D];
    const TaskParams& T = A.t;
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * STACK_WARPS + wi;
    if (env >= T.n_envs) return;
    const int D = T.D, L = T.n_lw;
    int32_t* w5 = A.p.env5 + (long long)env * ENV5_WORDS;
    const int mode = w5[W5_STACK_MODE];
    if (mode == STACK_KEEP) return;
    float* obs = A.obs_lidar + (long long)env * N_STACK * 3 * N_CELLS;
    int32_t* prev = A.p.stack_prev + (long long)env * 5 * D;
    const int prev_n = w5[W5_PREV_N];
    for (int i = lane; i < prev_n; i += 32) {
        const int code = prev[i], sp = code / N_CELLS, c = code - sp * N_CELLS;
        float* o = obs + sp * 3 * N_CELLS + c;
        o[0] = 1.0f; o[N_CELLS] = 1.0f; o[2 * N_CELLS] = 1.0f;
    }
    __syncwarp();
    uint8_t* mask = A.obs_mask + (long long)env * N_STACK;
    if (mode == STACK_EMPTY) {                 // reset observation: the ring was wiped by the step-0 broadcast
        if (lane < N_STACK) mask[lane] = 0;
        if (lane == 0) w5[W5_PREV_N] = 0;
        return;
    }
    const int ag = w5[W5_AGENT];
    const int cur = A.p.env[(long long)env * ENV_WORDS + W_STEP];
    const uint32_t call = (uint32_t)(w5[W5_OBS_CALL] - 1);
    const double my_u = lane < 14 ? philox_uniform(T.k0, T.k1, T.env_offset + (uint32_t)env, STREAM_FUSE, 16u * call + (uint32_t)lane, (uint32_t)ag) : 0.0;
    auto u = [&](int i) { return __shfl_sync(0xffffffffu, my_u, i); };
    // candidates: wingmen still publishing (never disarmed in this episode), in slot order
    int cands[8], m = 0;
    for (int P = 0; P < L && P < 8; ++P)
        if (A.p.flagw[(long long)env * D + P] & F_ARMED) cands[m++] = P;
    const int n = 1 + (int)(u(0) * 4.0);
    const int k = n < m ? n : m;
    for (int i = 0; i < k; ++i) {             // random.sample: partial Fisher-Yates
        const int j = i + (int)(u(1 + i) * (double)(m - i));
        const int t = cands[i]; cands[i] = cands[j]; cands[j] = t;
    }
    // sphere sources: 0 = own, then the chosen snapshots that exist (age <= steps since the reset)
    int srcP[5], srcAge[5], n_src = 1;
    srcP[0] = ag; srcAge[0] = 0;
    for (int i = 0; i < k; ++i) {
        const int a = 1 + (int)(u(5 + i) * 9.0);
        if (cur - a < 0) continue;
        srcP[n_src] = cands[i]; srcAge[n_src] = a; ++n_src;
    }
    int order[N_STACK];
    for (int i = 0; i < N_STACK; ++i) order[i] = i;
    for (int kk = 0, i = N_STACK - 1; i > 0; --i, ++kk) {   // random.shuffle
        const int j = (int)(u(9 + kk) * (double)(i + 1));
        const int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    int dst_of[N_STACK];
    for (int dst = 0; dst < N_STACK; ++dst) {
        dst_of[order[dst]] = dst;
        if (lane == 0) mask[dst] = order[dst] < n_src ? 1 : 0;
    }
    const float* own_pose = A.p.ring_pose + (((long long)env * L + ag) * RING + cur % RING) * 8;
    const double opx = own_pose[0], opy = own_pose[1], opz = own_pose[2];
    const double oqx = own_pose[3], oqy = own_pose[4], oqz = own_pose[5], oqw = own_pose[6];
    const double radius = 2 * T.dome;
    int n_new = 0;
    for (int src = 0; src < n_src; ++src) {
        const int P = srcP[src], a = srcAge[src], s = cur - a;
        // ring entries: pose of step s + 1; features: own -> step s + 1, neighbour -> step s (none at s = 0)
        const long long e_pose = ((long long)env * L + P) * RING + (s + 1) % RING;
        const long long e_feat = (src == 0 || P == ag) ? ((long long)env * L + ag) * RING + (src == 0 ? cur : s + 1) % RING
                                                       : ((long long)env * L + P) * RING + s % RING;
        const bool has_feats = src == 0 || P == ag || s >= 1;
        const float* np_ = A.p.ring_pose + e_pose * 8;
        double r00 = 1, r01 = 0, r02 = 0, r10 = 0, r11 = 1, r12 = 0, r20 = 0, r21 = 0, r22 = 1, npx = 0, npy = 0, npz = 0;
        if (src > 0) {
            npx = np_[0]; npy = np_[1]; npz = np_[2];
            const double x = np_[3], y = np_[4], z = np_[5], w = np_[6];
            r00 = 1 - 2 * (y * y + z * z); r01 = 2 * (x * y - w * z); r02 = 2 * (x * z + w * y);
            r10 = 2 * (x * y + w * z); r11 = 1 - 2 * (x * x + z * z); r12 = 2 * (y * z - w * x);
            r20 = 2 * (x * z - w * y); r21 = 2 * (y * z + w * x); r22 = 1 - 2 * (x * x + y * y);
        }
        for (int d = lane; d < D; d += 32) {
            int cell = -1; double rn = 1.0;
            const int meta = has_feats ? A.p.ring_meta[e_feat * D + d] : -1;
            if (meta >= 0) {
                const double* f = A.p.ring_feat + (e_feat * D + d) * 3;
                if (src == 0) { cell = meta & 0xffff; rn = f[0]; }
                else if (d != ag) {            // transform_features skips the observer's own echo
                    const double Rr = f[0] * radius;
                    double st, ct, sp, cp;
                    sincos(f[1], &st, &ct); sincos(f[2], &sp, &cp);
                    const double cx = Rr * st * cp, cy = Rr * st * sp, cz = Rr * ct;
                    const double gx = (r00 * cx + r01 * cy + r02 * cz) + npx, gy = (r10 * cx + r11 * cy + r12 * cz) + npy,
                                 gz = (r20 * cx + r21 * cy + r22 * cz) + npz;
                    const LidarHit h = lidar_project_one(0, radius, opx, opy, opz, oqx, oqy, oqz, oqw, gx, gy, gz);
                    cell = h.cell; rn = h.rn;
                }
            }
            s_cell[wi][d] = cell; s_rn[wi][d] = rn;
        }
        __syncwarp();
        const int dst = dst_of[src];
        float* o = obs + dst * 3 * N_CELLS;
        const float tval = src == 0 ? 0.1f : (float)fmin(fmax((double)a / RING, 0.0), 1.0);
        for (int d0 = 0; d0 < D; d0 += 32) {
            const int d = d0 + lane;
            bool win = false; int c = -1;
            if (d < D && (c = s_cell[wi][d]) >= 0) {
                if (src == 0) win = true;      // kept features are one per cell already
                else {
                    float curv = 1.0f; int wj = -1;
                    for (int j = 0; j < D; ++j) {
                        if (s_cell[wi][j] != c) continue;
                        const double rj = s_rn[wi][j];
                        const bool take = curv < 1.0f ? (rj > (double)curv) : true;
                        if (take) { curv = (float)rj; wj = j; }
                    }
                    win = wj == d;
                }
            }
            if (win) {
                const int meta = A.p.ring_meta[e_feat * D + d];
                o[c] = (float)s_rn[wi][d];
                o[N_CELLS + c] = (float)((double)(meta >> 16) / 5.0);
                o[2 * N_CELLS + c] = tval;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, win);
            if (win) prev[n_new + __popc(bal & ((1u << lane) - 1))] = dst * N_CELLS + c;
            n_new += __popc(bal);
        }
        __syncwarp();
    }
    if (lane == 0) w5[W5_PREV_N] = n_new;
}

}  // namespace dc
