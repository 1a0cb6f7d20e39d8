// dronechase_b200 -- the fused stage03 env-step kernel.
//
// One launch == Env.step for every env of the shard (exp02_vFinal_environment.py:155-188):
//   P1  scripted pilots        Task.on_step_start   exp02_vFinal_task.py:231-242,275-282
//   P2  16 physics substeps    advance_step         exp02_vFinal_environment.py:179-188
//   P3  engagement/reward/termination/info/obs vector/waves/auto-reset
//                              Task.on_step_middle  exp02_vFinal_task.py:284-318
//                              Task.on_step_end     exp02_vFinal_task.py:320-332
//   P4  projection LiDAR       compute_observation  exp02_vFinal_environment.py:206-234
//   P5  write-back of the struct-of-arrays drone state
// Thread map: a block owns EPB consecutive envs; thread t is drone (t % D) of local env (t / D);
// slot 0 of each env is the RL agent and doubles as that env's logic thread in P3.  Drone state
// lives in registers from the coalesced 16-byte loads of P0 to the stores of P5; everything that
// crosses drones goes through shared memory.
#pragma once
#include "common.cuh"
#include "lidar.cuh"
#include "quad_dynamics.cuh"

namespace dc {

enum { MODE_STEP = 0, MODE_RESET = 1 };
enum { NAV_WAIT = 0, NAV_WINGMAN = 1, NAV_BUILDING = 2 };
// per-drone flag word (state quad 0 .w)
enum { F_ARMED = 1, F_OFF = 2, F_NAV_SHIFT = 2 };
// per-drone event word built by the env's logic thread
enum { EV_LIVE = 1, EV_OFF = 2, EV_MID = 4, EV_ZEROED = 8, EV_REPLACED = 16, EV_REARMED = 32 };
// env scalar words
enum { W_STEP = 0, W_MAX_STEP, W_ROUND, W_AGENT_KILLS, W_ALLIES_KILLS, W_DEADS, W_BUILDING, W_HIT_CTR,
       W_SPAWN_CTR, W_PHYS_CTR, W_LAST_CLOSEST_LO, W_LAST_CLOSEST_HI, W_EP_RETURN, W_EP_STEPS, W_INIT, W_SPARE };

struct TaskParams {
    int n_envs, n_lw, n_lm, D;
    int munition, step_increment, max_step, initial_round, substeps;
    int lm_nav, ally_mode, reward, lidar, fixed_lw_spawn, auto_reset;
    uint32_t env_offset, k0, k1;
    double dome, born, lw_spawn, expl, shoot, cooldown, fire_p, lm_speed, bt_speed, ally_stop, vel_bonus;
    double building[3];
};

template <typename R> struct StepArgs {
    TaskParams t;
    QuadParams<R> q;
    V4<R>* state;            // [STATE_QUADS][E*D]
    int32_t* env;            // [E][ENV_WORDS]
    double* lw_init;         // [E][n_lw][3]
    const float* actions;
    float* obs_lidar; float* obs_inertial; float* obs_last_action;
    float* reward; uint8_t* done; int32_t* info; int32_t* lidar_ids;
    float* term_inertial; float* term_last_action; double* stats;
    const uint8_t* reset_mask;
    int epb;                 // envs per block
};

__device__ __forceinline__ double norm3(double x, double y, double z) { return sqrt(x * x + y * y + z * z); }

// Shared-memory view of one block.
template <typename R> struct Smem {
    R* ipos;        // [NS][3] imu position (offsets snapshot before the dynamics, fresh imu after)
    R* newpos;      // [NS][3]
    R* last;        // [NS] last_fired_step
    int* flags;     // [NS] F_ARMED | F_OFF (snapshot) -- read-only during P1
    int* ev;        // [NS] EV_*
    int* ammo;      // [NS]
    int* cell;      // [NS]
    double* rn;     // [NS]
    R* aquat;       // [EPB][4] agent imu quaternion
    int* envflag;   // [EPB] bit0: rewrite this env's sphere, bit1: nav reset, bit2: offsets refreshed
};

template <typename R>
__device__ __forceinline__ Smem<R> carve_smem(unsigned char* base, int ns, int epb) {
    Smem<R> s;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base + off; off += (bytes + 15) & ~size_t(15); return p; };
    s.rn = (double*)take(sizeof(double) * ns);
    s.ipos = (R*)take(sizeof(R) * 3 * ns);
    s.newpos = (R*)take(sizeof(R) * 3 * ns);
    s.last = (R*)take(sizeof(R) * ns);
    s.aquat = (R*)take(sizeof(R) * 4 * epb);
    s.flags = (int*)take(sizeof(int) * ns);
    s.ev = (int*)take(sizeof(int) * ns);
    s.ammo = (int*)take(sizeof(int) * ns);
    s.cell = (int*)take(sizeof(int) * ns);
    s.envflag = (int*)take(sizeof(int) * epb);
    return s;
}

inline size_t smem_bytes(int ns, int epb, size_t sizeofR) {
    auto up = [](size_t b) { return (b + 15) & ~size_t(15); };
    return up(8 * ns) + 2 * up(sizeofR * 3 * ns) + up(sizeofR * ns) + up(sizeofR * 4 * epb) + 4 * up(4 * ns) + up(4 * epb);
}

// ------------------------------------------------------------------------------------------------
// Logic-thread helpers.  `b` = index of the env's slot 0 inside the block's shared arrays.
// ------------------------------------------------------------------------------------------------
template <typename R> struct EnvCtx {
    const TaskParams& T;
    Smem<R>& S;
    int b;                   // base slot of this env in shared memory
    uint32_t env_id;         // global env index (Philox counter word)
    int32_t* w;              // env scalar words (registers/local copy)
    __device__ EnvCtx(const TaskParams& t, Smem<R>& s, int base, uint32_t id, int32_t* words)
        : T(t), S(s), b(base), env_id(id), w(words) {}

    __device__ void disarm(int d) {                       // Quadcopter.disarm quadcopter.py:461-478
        S.ev[b + d] = (S.ev[b + d] & ~EV_LIVE) | EV_ZEROED;
    }
    __device__ void arm(int d) {                          // Quadcopter.arm quadcopter.py:445-459
        S.ev[b + d] |= EV_LIVE | EV_REARMED;
        if (d < T.n_lw) { S.ammo[b + d] = T.munition; S.last[b + d] = (R)(-T.cooldown); }
    }
    __device__ void replace(int d, double x, double y, double z) {   // quadcopter.py:433-439
        S.ev[b + d] |= EV_REPLACED;
        S.newpos[3 * (b + d) + 0] = (R)x; S.newpos[3 * (b + d) + 1] = (R)y; S.newpos[3 * (b + d) + 2] = (R)z;
    }
    __device__ bool live(int d) const { return S.ev[b + d] & EV_LIVE; }
    __device__ bool off(int d) const { return S.ev[b + d] & EV_OFF; }
    __device__ double spawn_u(uint32_t idx) const {
        return philox_uniform(T.k0, T.k1, env_id, STREAM_SPAWN, idx);
    }
    // generate_positions(n, r)[i]  exp02_vFinal_task.py:583-607 (thetas drawn first, then phis)
    __device__ void gen_position(uint32_t base, int n, int i, double r, double* out) const {
        const double PI = 3.141592653589793;
        const double theta = 0.0 + (PI - 0.0) * spawn_u(base + i);
        const double min_z = 4.0;
        const double lower = fmin(min_z, r);
        const double lo = (r >= min_z) ? acos(lower / r) : 0.0;
        const double phi = lo + (PI / 2 - lo) * spawn_u(base + n + i);
        out[0] = r * sin(phi) * cos(theta);
        out[1] = r * sin(phi) * sin(theta);
        out[2] = r * cos(phi);
    }
    // setup_round(k)  exp02_vFinal_task.py:179-195
    __device__ void setup_round(int k) {
        for (int i = 0; i < T.n_lm; ++i) disarm(T.n_lw + i);
        const uint32_t base = (uint32_t)w[W_SPAWN_CTR];
        for (int i = 0; i < k; ++i) {
            double p[3];
            gen_position(base, k, i, T.born, p);
            replace(T.n_lw + i, p[0], p[1], p[2]);
            arm(T.n_lw + i);
        }
        w[W_SPAWN_CTR] += 2 * k;
    }
    // OffsetHandler.on_episode_start: snapshot := live set
    __device__ void refresh_offsets() {
        for (int d = 0; d < T.D; ++d) {
            int e = S.ev[b + d];
            S.ev[b + d] = (e & EV_LIVE) ? (e | EV_OFF) : (e & ~EV_OFF);
        }
    }
    // Task.on_episode_start  exp02_vFinal_task.py:258-267
    __device__ void episode_start(double* lw_init) {
        w[W_ROUND] = T.initial_round;
        setup_round(T.initial_round);
        for (int j = 0; j < T.n_lw; ++j) arm(j);
        const uint32_t base = (uint32_t)w[W_SPAWN_CTR];
        for (int j = 0; j < T.n_lw; ++j) {
            double p[3];
            if (T.fixed_lw_spawn) { p[0] = lw_init[3 * j]; p[1] = lw_init[3 * j + 1]; p[2] = lw_init[3 * j + 2]; }
            else gen_position(base, T.n_lw, j, T.lw_spawn, p);
            replace(j, p[0], p[1], p[2]);
        }
        if (!T.fixed_lw_spawn) w[W_SPAWN_CTR] += 2 * T.n_lw;
    }
    // Env.__init__: Task.on_env_init + on_episode_start  exp02_vFinal_environment.py:62-63, task :248-252,622-646
    __device__ void env_init(double* lw_init) {
        uint32_t base = (uint32_t)w[W_SPAWN_CTR];
        for (int i = 0; i < T.n_lm; ++i) {
            double p[3];
            gen_position(base, T.n_lm, i, T.born, p);
            replace(T.n_lw + i, p[0], p[1], p[2]);
        }
        w[W_SPAWN_CTR] += 2 * T.n_lm;
        base = (uint32_t)w[W_SPAWN_CTR];
        for (int j = 0; j < T.n_lw; ++j) {
            double p[3];
            gen_position(base, T.n_lw, j, T.lw_spawn, p);
            lw_init[3 * j] = p[0]; lw_init[3 * j + 1] = p[1]; lw_init[3 * j + 2] = p[2];
            replace(j, p[0], p[1], p[2]);
        }
        w[W_SPAWN_CTR] += 2 * T.n_lw;
        episode_start(lw_init);
        w[W_INIT] = 1;
    }
    // Env.reset -> Task.on_reset  exp02_vFinal_environment.py:133-151, task :254-273
    __device__ void reset_env(double* lw_init) {
        w[W_STEP] = 0; w[W_MAX_STEP] = T.max_step;
        w[W_AGENT_KILLS] = w[W_ALLIES_KILLS] = w[W_DEADS] = 0; w[W_BUILDING] = 1;
        set_last_closest(T.dome);
        w[W_EP_RETURN] = __float_as_int(0.0f); w[W_EP_STEPS] = 0;
        for (int d = 0; d < T.D; ++d) disarm(d);
        episode_start(lw_init);
        refresh_offsets();
        S.envflag[b_env()] |= 2 | 4;
    }
    __device__ int b_env() const { return b / T.D; }
    __device__ void set_last_closest(double v) {
        long long bits = __double_as_longlong(v);
        w[W_LAST_CLOSEST_LO] = (int32_t)(bits & 0xffffffffLL); w[W_LAST_CLOSEST_HI] = (int32_t)(bits >> 32);
    }
    __device__ double last_closest() const {
        long long bits = ((long long)w[W_LAST_CLOSEST_HI] << 32) | (unsigned int)w[W_LAST_CLOSEST_LO];
        return __longlong_as_double(bits);
    }
    __device__ double pos(int d, int k) const { return (double)S.ipos[3 * (b + d) + k]; }
    __device__ double dist(int a, int c) const {
        return norm3(pos(a, 0) - pos(c, 0), pos(a, 1) - pos(c, 1), pos(a, 2) - pos(c, 2));
    }
    // Gun.is_available gun.py:56-75 (current_step == env step after the broadcast)
    __device__ bool gun_available(int j) const {
        if (S.ammo[b + j] <= 0) return true;
        return T.cooldown <= (double)w[W_STEP] - (double)S.last[b + j];
    }
    // nearest snapshot invader of pursuer j with d < thr (identify_invaders_in_range(...)[j][0]
    // offsets_handler.py:283-309: ascending stable sort -> first index wins ties); -1 if none
    __device__ int nearest_in_range(int j, double thr) const {
        int best = -1; double bd = 0.0;
        for (int i = T.n_lw; i < T.D; ++i) {
            if (!off(i)) continue;
            const double d = dist(j, i);
            if (d < thr && (best < 0 || d < bd)) { best = i; bd = d; }
        }
        return best;
    }
    // identify_closest_invader(src) offsets_handler.py:256-281 (np.argmin: first index on ties)
    __device__ int nearest_invader(int src) const {
        int best = -1; double bd = 0.0;
        for (int i = T.n_lw; i < T.D; ++i) {
            if (!off(i)) continue;
            const double d = dist(src, i);
            if (best < 0 || d < bd) { best = i; bd = d; }
        }
        return best;
    }
    __device__ int count_outside_dome(int lo, int hi) const {
        int n = 0;
        for (int d = lo; d < hi; ++d)
            if (off(d) && norm3(pos(d, 0), pos(d, 1), pos(d, 2)) > T.dome) ++n;
        return n;
    }
};

// ------------------------------------------------------------------------------------------------
template <typename R, int MODE, bool NOISE>
__global__ void __launch_bounds__(256) stage03_kernel(const StepArgs<R> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TaskParams& T = A.t;
    const int D = T.D, EPB = A.epb, NS = EPB * D;
    Smem<R> S = carve_smem<R>(smem_raw, NS, EPB);
    const int tid = threadIdx.x;
    const int env0 = blockIdx.x * EPB;
    const int nenv = min(EPB, T.n_envs - env0);
    const int le = tid / D, d = tid - le * D;
    const bool has_drone = tid < nenv * D;
    const int env = env0 + le;
    const long long slot = (long long)env * D + d;
    const long long stride = (long long)T.n_envs * D;
    const bool is_lw = d < T.n_lw;
    const bool is_logic = has_drone && d == 0;

    // ---- P0: coalesced loads -------------------------------------------------------------------
    Drone<R> s; R ipx = 0, ipy = 0, ipz = 0, fox = 0, foy = 0, foz = 0, lastf = 0;
    int flags = 0, ammo = 0;
    int32_t w[ENV_WORDS];
    if (has_drone) {
        const V4<R>* st = A.state + slot;
        V4<R> v = ld4(st); s.px = v.x; s.py = v.y; s.pz = v.z; flags = (int)v.w;
        v = ld4(st + stride); s.qx = v.x; s.qy = v.y; s.qz = v.z; s.qw = v.w;
        v = ld4(st + 2 * stride); s.vx = v.x; s.vy = v.y; s.vz = v.z; lastf = v.w;
        v = ld4(st + 3 * stride); s.wx = v.x; s.wy = v.y; s.wz = v.z; ammo = (int)v.w;
        v = ld4(st + 4 * stride); s.thr[0] = v.x; s.thr[1] = v.y; s.thr[2] = v.z; s.thr[3] = v.w;
#pragma unroll
        for (int k = 0; k < 5; ++k) {       // PID words 0..19 (20..23 belong to mode 7 only)
            v = ld4(st + (5 + k) * stride);
            s.pid[4 * k] = v.x; s.pid[4 * k + 1] = v.y; s.pid[4 * k + 2] = v.z; s.pid[4 * k + 3] = v.w;
        }
        v = ld4(st + 11 * stride); ipx = v.x; ipy = v.y; ipz = v.z;
        if (is_lw) { v = ld4(st + 12 * stride); fox = v.x; foy = v.y; foz = v.z; }
        S.ipos[3 * tid] = ipx; S.ipos[3 * tid + 1] = ipy; S.ipos[3 * tid + 2] = ipz;
        S.flags[tid] = flags & 3;
        S.ammo[tid] = ammo; S.last[tid] = lastf;
        if (is_logic) {
            const int4* wp = reinterpret_cast<const int4*>(A.env + (long long)env * ENV_WORDS);
#pragma unroll
            for (int k = 0; k < 4; ++k) { int4 t = wp[k]; w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w; }
            S.envflag[le] = 0;
        }
    }
    __syncthreads();

    Imu<R> imu;
    imu.px = ipx; imu.py = ipy; imu.pz = ipz; imu.roll = imu.pitch = imu.yaw = 0;
    imu.ub = imu.vb = imu.wb = imu.p = imu.q = imu.r = 0; imu.qx = imu.qy = imu.qz = 0; imu.qw = 1;
    int nav = (flags >> F_NAV_SHIFT) & 3;
    const int b = le * D;            // this env's base slot in shared memory
    float act[4] = {0.f, 0.f, 0.f, 0.f};

    if (MODE == MODE_STEP) {
        // ---- P1: scripted pilots + RL action -> mode-6 setpoint --------------------------------
        R sp[4] = {0, 0, 0, 0};
        const bool armed = has_drone && (flags & F_ARMED);
        if (has_drone) {
            double cmd[4] = {0, 0, 0, 0};
            bool driven = false;
            if (d == 0) {
                const float4 a = reinterpret_cast<const float4*>(A.actions)[env];
                act[0] = a.x; act[1] = a.y; act[2] = a.z; act[3] = a.w;
                cmd[0] = a.x; cmd[1] = a.y; cmd[2] = a.z; cmd[3] = a.w; driven = true;
            } else if (armed && !is_lw) {
                // KamikazeNavigator.update: check_transition, then execute the state fetched BEFORE it
                const double mx = ipx, my = ipy, mz = ipz;
                bool any_lw = false;
                for (int j = 0; j < T.n_lw; ++j) any_lw |= (S.flags[b + j] & F_OFF) != 0;
                auto path_clear = [&](double degrees) {
                    if (T.lm_nav == 0) return false;             // air_combat_only :87-96 always False
                    const double abx = T.building[0] - mx, aby = T.building[1] - my, abz = T.building[2] - mz;
                    const double nab = norm3(abx, aby, abz);
                    for (int j = 0; j < T.n_lw; ++j) {
                        if (!(S.flags[b + j] & F_OFF)) continue;
                        const double apx = (double)S.ipos[3 * (b + j)] - mx, apy = (double)S.ipos[3 * (b + j) + 1] - my,
                                     apz = (double)S.ipos[3 * (b + j) + 2] - mz;
                        const double nap = norm3(apx, apy, apz);
                        if (nap > nab) continue;
                        const double ang = acos((apx * abx + apy * aby + apz * abz) / (nap * nab)) * (180.0 / 3.141592653589793);
                        if (ang <= degrees / 2) return false;     // a wingman sits inside the cone
                    }
                    return true;
                };
                double tx = 0, ty = 0, tz = 0; bool moving = false;
                if (nav == NAV_WAIT) {
                    if (path_clear(60.0)) nav = NAV_BUILDING;
                    else if (any_lw) nav = NAV_WINGMAN;
                } else if (nav == NAV_WINGMAN) {
                    if (!any_lw) nav = NAV_BUILDING;
                    int best = -1; double bd = 0;
                    for (int j = 0; j < T.n_lw; ++j) {
                        if (!(S.flags[b + j] & F_OFF)) continue;
                        const double dd = norm3((double)S.ipos[3 * (b + j)] - mx, (double)S.ipos[3 * (b + j) + 1] - my,
                                                (double)S.ipos[3 * (b + j) + 2] - mz);
                        if (best < 0 || dd < bd) { best = j; bd = dd; }
                    }
                    if (best >= 0) { tx = S.ipos[3 * (b + best)]; ty = S.ipos[3 * (b + best) + 1]; tz = S.ipos[3 * (b + best) + 2]; }
                    moving = true;
                } else {
                    if (!path_clear(45.0)) nav = NAV_WINGMAN;
                    tx = T.building[0]; ty = T.building[1]; tz = T.building[2]; moving = true;
                }
                if (moving) {
                    const double vx = tx - mx, vy = ty - my, vz = tz - mz, n = norm3(vx, vy, vz);
                    if (n > 0) { cmd[0] = vx / n; cmd[1] = vy / n; cmd[2] = vz / n; } else { cmd[0] = vx; cmd[1] = vy; cmd[2] = vz; }
                }
                cmd[3] = T.lm_speed; driven = true;
            } else if (armed && is_lw) {
                // drive_loyalwingmen: get_armed_pursuers()[1:]
                int armed_before = 0;
                for (int j = 0; j < d; ++j) armed_before += (S.flags[b + j] & F_ARMED) ? 1 : 0;
                if (armed_before >= 1) {
                    if (T.ally_mode == 1) { cmd[3] = T.ally_stop; }
                    else {
                        // LoyalWingmanBehaviorTree: gun available (or empty) -> chase, else formation
                        const int cur_step = A.env[(long long)env * ENV_WORDS + W_STEP];
                        const bool avail = ammo <= 0 || T.cooldown <= (double)cur_step - (double)lastf;
                        double tx = ipx, ty = ipy, tz = ipz;
                        if (avail) {
                            int best = -1; double bd = 0;
                            for (int i = T.n_lw; i < D; ++i) {
                                if (!(S.flags[b + i] & F_OFF)) continue;
                                const double dd = norm3((double)S.ipos[3 * (b + i)] - (double)ipx, (double)S.ipos[3 * (b + i) + 1] - (double)ipy,
                                                        (double)S.ipos[3 * (b + i) + 2] - (double)ipz);
                                if (best < 0 || dd < bd) { best = i; bd = dd; }
                            }
                            if (best >= 0) { tx = S.ipos[3 * (b + best)]; ty = S.ipos[3 * (b + best) + 1]; tz = S.ipos[3 * (b + best) + 2]; }
                        } else { tx = fox; ty = foy; tz = foz; }
                        const double vx = tx - (double)ipx, vy = ty - (double)ipy, vz = tz - (double)ipz, n = norm3(vx, vy, vz);
                        if (n > 0) { cmd[0] = vx / n; cmd[1] = vy / n; cmd[2] = vz / n; } else { cmd[0] = vx; cmd[1] = vy; cmd[2] = vz; }
                        cmd[3] = T.bt_speed;
                    }
                    driven = true;
                }
            }
            if (driven) {                                   // convert_command_to_setpoint quadcopter.py:379-396
                const double n = norm3(cmd[0], cmd[1], cmd[2]);
                const double dn = n > 0 ? n : 1.0;
                sp[0] = (R)(cmd[3] * (cmd[0] / dn)); sp[1] = (R)(cmd[3] * (cmd[1] / dn));
                sp[2] = 0; sp[3] = (R)(cmd[3] * (cmd[2] / dn));
            }
        }
        __syncthreads();      // all P1 reads of S.ipos/S.flags done before P2 publishes fresh imu data

        // ---- P2: physics substeps, state in registers ------------------------------------------
        if (armed) {
            const uint32_t env_id = T.env_offset + (uint32_t)env;
            const uint32_t phys0 = (uint32_t)A.env[(long long)env * ENV_WORDS + W_PHYS_CTR];
            for (int k = 0; k < T.substeps; ++k)
                quad_substep<R, NOISE>(s, sp, A.q, imu, T.k0, T.k1, env_id, (uint32_t)d, phys0 + (uint32_t)k);
            ipx = imu.px; ipy = imu.py; ipz = imu.pz;
        }
        if (has_drone) {
            S.ipos[3 * tid] = ipx; S.ipos[3 * tid + 1] = ipy; S.ipos[3 * tid + 2] = ipz;
            S.ev[tid] = (flags & F_ARMED) ? (EV_LIVE | EV_OFF) : 0;   // on_middle_step: snapshot := armed set
            if (d == 0) { S.aquat[4 * le] = imu.qx; S.aquat[4 * le + 1] = imu.qy; S.aquat[4 * le + 2] = imu.qz; S.aquat[4 * le + 3] = imu.qw; }
        }
    } else {
        if (has_drone) S.ev[tid] = ((flags & F_ARMED) ? EV_LIVE : 0) | ((flags & F_OFF) ? EV_OFF : 0);
    }
    __syncthreads();

    // ---- P3: per-env game logic on the agent's thread ---------------------------------------------
    if (is_logic) {
        EnvCtx<R> C(T, S, b, T.env_offset + (uint32_t)env, w);
        double* lw_init = A.lw_init + (long long)env * T.n_lw * 3;
        float inertial[15];
        auto gun_state = [&](float* g) {                  // Gun.get_state gun.py:101-113
            const double wait = fmax(T.cooldown - ((double)w[W_STEP] - (double)S.last[b]), 0.0);
            const int mx = T.munition > 0 ? T.munition : 1;
            g[0] = (float)((double)S.ammo[b] / (double)mx);
            g[1] = (float)(wait / T.cooldown);
            g[2] = C.gun_available(0) ? 1.f : 0.f;
        };
        if (MODE == MODE_STEP) {
            w[W_STEP] += 1; w[W_PHYS_CTR] += T.substeps; w[W_EP_STEPS] += 1;
            if (T.reward == 1) {                           // update_building_life (exp02_v2_full_task.py)
                int cnt = 0;
                for (int i = T.n_lw; i < D; ++i)
                    if (C.off(i) && norm3(C.pos(i, 0), C.pos(i, 1), C.pos(i, 2)) < 0.2) ++cnt;
                w[W_BUILDING] = max(w[W_BUILDING] - cnt, 0);
            }
            // process_shoot_range_invaders :391-412 -> shoot_by_ids -> Gun.shoot
            int agent_shots = 0, ally_shots = 0;
            for (int j = 0; j < T.n_lw; ++j) {
                if (!C.off(j)) continue;
                const int tgt = C.nearest_in_range(j, T.shoot);
                if (tgt < 0) continue;
                if (!(C.gun_available(j) && S.ammo[b + j] > 0)) continue;
                S.ammo[b + j] -= 1; S.last[b + j] = (R)w[W_STEP];
                const double u = philox_uniform(T.k0, T.k1, C.env_id, STREAM_HIT, (uint32_t)w[W_HIT_CTR]);
                w[W_HIT_CTR] += 1;
                if (u < T.fire_p) { C.disarm(tgt); if (j == 0) ++agent_shots; else ++ally_shots; }
            }
            // process_explosion_range_invaders :358-389 (same, now stale, distance matrix)
            int exploded = 0, ally_suicide = 0, agent_suicide = 0;
            for (int j = 0; j < T.n_lw; ++j) {
                if (!C.off(j)) continue;
                const int tgt = C.nearest_in_range(j, T.expl);
                if (tgt < 0) continue;
                C.disarm(j); C.disarm(tgt);
                if (T.reward == 1) ++exploded;
                else if (S.ammo[b + j] == 0 && j == 0) ++agent_suicide;
                else if (S.ammo[b + j] == 0) ++ally_suicide;
                else ++exploded;
            }
            w[W_AGENT_KILLS] += agent_shots; w[W_ALLIES_KILLS] += ally_shots; w[W_DEADS] += exploded;
            for (int i = T.n_lw; i < D; ++i)               // process_invaders_in_origin :656-659
                if (C.off(i) && norm3(C.pos(i, 0), C.pos(i, 1), C.pos(i, 2)) < 0.2) C.disarm(i);

            // ---- reward ----
            float g[3]; gun_state(g);
            const double apx = imu.px, apy = imu.py, apz = imu.pz;
            const int lw_out = C.count_outside_dome(0, T.n_lw);
            double reward;
            if (T.reward == 0) {                           // exp02_vFinal_task.py:422-514
                double bonus = 0, penalty = 0, score;
                const double munition = (double)S.ammo[b] / (double)(T.munition > 0 ? T.munition : 1);
                const double reload = fmax(T.cooldown - ((double)w[W_STEP] - (double)S.last[b]), 0.0) / T.cooldown;
                const bool avail = C.gun_available(0);
                int src = -1;
                if (C.off(0)) {
                    int n_all = 0;
                    for (int j = 0; j < T.n_lw; ++j) n_all += C.off(j) ? 1 : 0;
                    if (n_all <= 1) src = 0;
                    else {
                        double bd = 0;
                        for (int j = 1; j < T.n_lw; ++j) {
                            if (!C.off(j)) continue;
                            const double dd = C.dist(j, 0);
                            if (src < 0 || dd < bd) { src = j; bd = dd; }
                        }
                    }
                }
                const int target = src >= 0 ? C.nearest_invader(src) : -1;
                double tpx = 0, tpy = 0, tpz = 0;
                if (target >= 0) { tpx = C.pos(target, 0); tpy = C.pos(target, 1); tpz = C.pos(target, 2); }
                const double current = norm3(apx - tpx, apy - tpy, apz - tpz);
                if (0.01 < C.last_closest() - current && (avail || munition == 0.0))
                    bonus += T.vel_bonus * norm3((double)imu.ub, (double)imu.vb, (double)imu.wb);
                C.set_last_closest(current);
                score = (avail || munition == 0.0) ? -current : current * (2 * reload - 1);
                if (agent_shots > 0 || agent_suicide > 0) bonus += (agent_shots + agent_suicide) * 1000.0;
                if (ally_shots > 0 || ally_suicide > 0) bonus += 0.5 * (ally_shots + ally_suicide) * 1000.0;
                else if (exploded > 0) penalty += 1000.0 * exploded;
                if (apz < -5.0) penalty += (-5.0 - apz) / (-5.0 + 6.0) * 1000.0;
                if (lw_out > 0) penalty += 1000.0;
                const double d0 = norm3(apx, apy, apz);
                if (d0 > T.born - 2) penalty += d0 - T.born - 2;
                reward = score + bonus - penalty;
            } else {                                       // exp02_v2_full_task.py compute_reward
                double bonus = 0, penalty = 0;
                const int shots = agent_shots + ally_shots;
                const double kills = (double)(w[W_AGENT_KILLS] + w[W_ALLIES_KILLS]);
                if (shots > 0) bonus += (shots + kills / 10) * 1000.0;
                if (S.ammo[b] == 0 && exploded > 0) bonus += (shots + kills / 10) * 1000.0;
                else if (exploded > 0) penalty += 1000.0 * exploded;
                if (apz < 0.01) penalty += 1000.0;
                if (lw_out > 0) penalty += 1000.0;
                if (w[W_BUILDING] < 1) penalty += 1000.0 * (1 - w[W_BUILDING]);
                const double d0 = norm3(apx, apy, apz);
                if (d0 > T.born) penalty += d0 - T.born;
                reward = 0 + bonus - penalty;
            }
            if (agent_shots + ally_shots > 0) w[W_MAX_STEP] += T.step_increment;   // increment_max_step :149-152

            // ---- termination :516-568 ----
            bool lm_alive = false, lw_alive = false;
            for (int i = T.n_lw; i < D; ++i) lm_alive |= C.live(i);
            for (int j = 0; j < T.n_lw; ++j) lw_alive |= C.live(j);
            const bool all_over = !lm_alive && w[W_ROUND] >= T.n_lm;
            bool done = w[W_STEP] > w[W_MAX_STEP];
            done |= all_over;
            if (T.reward == 1) done |= w[W_BUILDING] <= 0;
            done |= lw_out > 0;
            done |= C.count_outside_dome(T.n_lw, D) > 0;
            done |= !lw_alive;
            done |= !C.live(0);
            done |= apz < (T.reward == 1 ? 0.01 : -5.99);

            w[W_EP_RETURN] = __float_as_int(__int_as_float(w[W_EP_RETURN]) + (float)reward);
            A.reward[env] = (float)reward;
            A.done[env] = done ? 1 : 0;
            int32_t* info = A.info + (long long)env * INFO_WORDS;
            reinterpret_cast<int4*>(info)[0] = make_int4(w[W_AGENT_KILLS], w[W_ALLIES_KILLS], w[W_DEADS], w[W_ROUND]);
            reinterpret_cast<int4*>(info)[1] = make_int4(w[W_BUILDING], w[W_STEP], w[W_MAX_STEP], w[W_EP_STEPS]);

            // ---- observation vector: normalize_inertial_data normalization.py:6-110 + gun_state ----
            const double PI = 3.141592653589793;
            const double max_speed = 1 * 10 * (1000.0 / 3600.0);
            auto nrm = [](double v, double sc) { return (float)fmin(fmax(v / sc, -1.0), 1.0); };
            inertial[0] = nrm(apx, T.dome); inertial[1] = nrm(apy, T.dome); inertial[2] = nrm(apz, T.dome);
            inertial[3] = nrm(imu.ub, max_speed); inertial[4] = nrm(imu.vb, max_speed); inertial[5] = nrm(imu.wb, max_speed);
            inertial[6] = nrm(imu.roll, PI); inertial[7] = nrm(imu.pitch, PI); inertial[8] = nrm(imu.yaw, PI);
            inertial[9] = nrm(imu.p, 2 * PI); inertial[10] = nrm(imu.q, 2 * PI); inertial[11] = nrm(imu.r, 2 * PI);
            inertial[12] = g[0]; inertial[13] = g[1]; inertial[14] = g[2];

            // LiDAR is rebuilt only while the agent is still a publisher (fused_lidar.py:160-166)
            for (int k = 0; k < D; ++k) if (S.ev[b + k] & EV_LIVE) S.ev[b + k] |= EV_MID;
            if (C.live(0)) S.envflag[le] |= 1;

            // ---- Task.on_step_end :320-332 + advance_round :154-174 ----
            if (!all_over && !lm_alive && lw_alive) {
                w[W_ROUND] += (w[W_ROUND] < T.n_lm) ? 1 : T.n_lm;
                C.setup_round(w[W_ROUND]);
                C.refresh_offsets();
                S.envflag[le] |= 2 | 4;
            }
            // ---- VecEnv auto-reset (SB3 DummyVecEnv.step_wait semantics) ----
            if (done && T.auto_reset) {
                if (A.term_inertial) for (int k = 0; k < 15; ++k) A.term_inertial[(long long)env * 15 + k] = inertial[k];
                if (A.term_last_action) reinterpret_cast<float4*>(A.term_last_action)[env] = make_float4(act[0], act[1], act[2], act[3]);
                if (A.stats) {
                    atomicAdd(A.stats + 0, 1.0); atomicAdd(A.stats + 1, (double)__int_as_float(w[W_EP_RETURN]));
                    atomicAdd(A.stats + 2, (double)w[W_EP_STEPS]); atomicAdd(A.stats + 3, (double)w[W_AGENT_KILLS]);
                    atomicAdd(A.stats + 4, (double)w[W_ALLIES_KILLS]); atomicAdd(A.stats + 5, (double)w[W_DEADS]);
                    atomicAdd(A.stats + 6, (double)w[W_ROUND]);
                }
                C.reset_env(lw_init);
                act[0] = act[1] = act[2] = act[3] = 0.f;
                inertial[0] = nrm(S.newpos[3 * b], T.dome); inertial[1] = nrm(S.newpos[3 * b + 1], T.dome);
                inertial[2] = nrm(S.newpos[3 * b + 2], T.dome);
                for (int k = 3; k < 12; ++k) inertial[k] = 0.f;
                gun_state(g); inertial[12] = g[0]; inertial[13] = g[1]; inertial[14] = g[2];
            }
        } else {
            // ---- MODE_RESET: Env.__init__ on first use, then Env.reset for the masked envs ----
            const bool masked = A.reset_mask == nullptr || A.reset_mask[env] != 0;
            const bool first = w[W_INIT] == 0;
            if (first) {
                for (int k = 0; k < ENV_WORDS; ++k) w[k] = 0;
                C.env_init(lw_init);
                S.envflag[le] |= 8;                         // first use: start from an empty sphere
            }
            if (masked || first) C.reset_env(lw_init);
            if (masked || first) {
                auto nrm = [](double v, double sc) { return (float)fmin(fmax(v / sc, -1.0), 1.0); };
                float g[3]; gun_state(g);
                inertial[0] = nrm(S.newpos[3 * b], T.dome); inertial[1] = nrm(S.newpos[3 * b + 1], T.dome);
                inertial[2] = nrm(S.newpos[3 * b + 2], T.dome);
                for (int k = 3; k < 12; ++k) inertial[k] = 0.f;
                inertial[12] = g[0]; inertial[13] = g[1]; inertial[14] = g[2];
                S.envflag[le] |= 16;                        // write the observation vector
            }
        }
        if (MODE == MODE_STEP || (S.envflag[le] & 16)) {
            float* oi = A.obs_inertial + (long long)env * 15;
            for (int k = 0; k < 15; ++k) oi[k] = inertial[k];
            reinterpret_cast<float4*>(A.obs_last_action)[env] = make_float4(act[0], act[1], act[2], act[3]);
        }
        int4* wp = reinterpret_cast<int4*>(A.env + (long long)env * ENV_WORDS);
#pragma unroll
        for (int k = 0; k < 4; ++k) wp[k] = make_int4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
    __syncthreads();

    // ---- P4: projection LiDAR of the agent (slot 0) over the entities alive after engagement --------
    const int ch = T.lidar == 0 ? 3 : 2;
    if (MODE == MODE_STEP) {
        if (has_drone) {
            int cell = -1; double rn = 1.0;
            if (d != 0 && (S.ev[tid] & EV_MID) && (S.envflag[le] & 1)) {
                LidarHit h;
                if (T.lidar == 0)      // float32 snapshot (perception_snapshot.py:91-110)
                    h = lidar_project_one(0, 2 * T.dome, (double)(float)S.ipos[3 * b], (double)(float)S.ipos[3 * b + 1],
                                          (double)(float)S.ipos[3 * b + 2], (double)(float)S.aquat[4 * le], (double)(float)S.aquat[4 * le + 1],
                                          (double)(float)S.aquat[4 * le + 2], (double)(float)S.aquat[4 * le + 3],
                                          (double)(float)ipx, (double)(float)ipy, (double)(float)ipz);
                else
                    h = lidar_project_one(1, 2 * T.dome, (double)S.ipos[3 * b], (double)S.ipos[3 * b + 1], (double)S.ipos[3 * b + 2],
                                          (double)S.aquat[4 * le], (double)S.aquat[4 * le + 1], (double)S.aquat[4 * le + 2],
                                          (double)S.aquat[4 * le + 3], (double)ipx, (double)ipy, (double)ipz);
                cell = h.cell; rn = h.rn;
            }
            S.cell[tid] = cell; S.rn[tid] = rn;
        }
        __syncthreads();
        const bool winner = has_drone && lidar_wins(T.lidar, d, D, S.cell + b, S.rn + b);
        // fill: every sphere that is rebuilt starts from all ones (LIDARSpec.empty_sphere angle_grid.py:89-99)
        {
            const int per_env = ch * N_CELLS;
            const long long f0 = (long long)env0 * per_env, f1 = f0 + (long long)nenv * per_env;
            float* base = A.obs_lidar;
            long long a0 = (f0 + 3) & ~3LL; if (a0 > f1) a0 = f1;
            const long long a1 = a0 + ((f1 - a0) & ~3LL);
            for (long long f = f0 + tid; f < a0; f += blockDim.x)
                if (S.envflag[(int)((f - f0) / per_env)] & 1) base[f] = 1.0f;
            for (long long f = a0 + 4LL * tid; f < a1; f += 4LL * blockDim.x) {
                const int e_lo = (int)((f - f0) / per_env), e_hi = (int)((f + 3 - f0) / per_env);
                const bool u_lo = S.envflag[e_lo] & 1, u_hi = S.envflag[e_hi] & 1;
                if (u_lo && u_hi) *reinterpret_cast<float4*>(base + f) = make_float4(1.f, 1.f, 1.f, 1.f);
                else if (u_lo || u_hi)
                    for (int k = 0; k < 4; ++k)
                        if (S.envflag[(int)((f + k - f0) / per_env)] & 1) base[f + k] = 1.0f;
            }
            for (long long f = a1 + tid; f < f1; f += blockDim.x)
                if (S.envflag[(int)((f - f0) / per_env)] & 1) base[f] = 1.0f;
            if (A.lidar_ids) {
                const long long g0 = (long long)env0 * N_CELLS, g1 = g0 + (long long)nenv * N_CELLS;
                for (long long f = g0 + tid; f < g1; f += blockDim.x)
                    A.lidar_ids[f] = -1;     // features = [] when the update is skipped (fused_lidar.py:165)
            }
        }
        __syncthreads();
        if (winner) {
            float* sph = A.obs_lidar + (long long)env * ch * N_CELLS;
            const int c = S.cell[tid];
            sph[c] = (float)S.rn[tid];
            sph[N_CELLS + c] = (float)((is_lw ? 3.0 : 1.0) / 5.0);      // EntityType value / 5
            if (ch == 3) sph[2 * N_CELLS + c] = 0.1f;                    // normalised age 1/10 (lidar_buffer.py:98-99)
            if (A.lidar_ids) A.lidar_ids[(long long)env * N_CELLS + c] = d;
        }
    } else {
        // first use of an env: empty sphere
        const int per_env = ch * N_CELLS;
        for (int e = 0; e < nenv; ++e) {
            if (!(S.envflag[e] & 8)) continue;
            float* sph = A.obs_lidar + (long long)(env0 + e) * per_env;
            for (int f = tid; f < per_env; f += blockDim.x) sph[f] = 1.0f;
            if (A.lidar_ids) for (int f = tid; f < N_CELLS; f += blockDim.x) A.lidar_ids[(long long)(env0 + e) * N_CELLS + f] = -1;
        }
    }

    // ---- P5: apply the env's events to the drone and store --------------------------------------
    if (has_drone) {
        const int ev = S.ev[tid];
        if (ev & EV_ZEROED) { s.vx = s.vy = s.vz = 0; s.wx = s.wy = s.wz = 0; s.thr[0] = s.thr[1] = s.thr[2] = s.thr[3] = 0; }
        if (ev & EV_REPLACED) {
            s.px = S.newpos[3 * tid]; s.py = S.newpos[3 * tid + 1]; s.pz = S.newpos[3 * tid + 2];
            s.qx = s.qy = s.qz = 0; s.qw = 1; s.vx = s.vy = s.vz = 0; s.wx = s.wy = s.wz = 0;
            fox = s.px; foy = s.py; foz = s.pz;
        }
        const bool live = ev & EV_LIVE;
        if (is_lw) { ammo = S.ammo[tid]; lastf = S.last[tid]; }
        else if (ev & EV_REARMED) { ammo = 10; lastf = (R)(-T.cooldown); }
        if (live && (ev & (EV_REARMED | EV_REPLACED))) { ipx = s.px; ipy = s.py; ipz = s.pz; }
        if (S.envflag[le] & 2) nav = NAV_WAIT;
        const int nf = (live ? F_ARMED : 0) | ((ev & EV_OFF) ? F_OFF : 0) | (nav << F_NAV_SHIFT);
        V4<R>* st = A.state + slot;
        st4(st, V4<R>{s.px, s.py, s.pz, (R)nf});
        st4(st + stride, V4<R>{s.qx, s.qy, s.qz, s.qw});
        st4(st + 2 * stride, V4<R>{s.vx, s.vy, s.vz, lastf});
        st4(st + 3 * stride, V4<R>{s.wx, s.wy, s.wz, (R)ammo});
        st4(st + 4 * stride, V4<R>{s.thr[0], s.thr[1], s.thr[2], s.thr[3]});
        if (MODE == MODE_STEP) {
#pragma unroll
            for (int k = 0; k < 5; ++k)
                st4(st + (5 + k) * stride, V4<R>{s.pid[4 * k], s.pid[4 * k + 1], s.pid[4 * k + 2], s.pid[4 * k + 3]});
        }
        st4(st + 11 * stride, V4<R>{ipx, ipy, ipz, (R)0});
        if (is_lw || (ev & EV_REPLACED)) st4(st + 12 * stride, V4<R>{fox, foy, foz, (R)0});
    }
}

}  // namespace dc
