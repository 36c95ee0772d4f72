// mp_batch.cu -- batched Micropolis engine on sm_100a: one warp per city, state resident in HBM as
// one contiguous block per city (mp_state.h EnvLayout), kernels + the mpb_* / mpb_env_* C-ABI of
// include/gymcity_b200.h.  The frame kernel runs 32 (16 / 8 for small batches) cities per CTA in frame
// lockstep; the other kernels (tools, reset, gym pre / post) one city per CTA.
//
// What this replaces in the reference: the C++ `Micropolis` engine object as the gym path drives
// it through SWIG (gym_city/envs/corecontrol.py:36-357) -- construction, clearMap/generateMap,
// toolDown, setFunds, simTick / tickEngine, and the per-cell getters used to read state back.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gymcity_b200.h"
#include "mp_env.h"
#include "mp_gym.h"
#include "mp_snapshot_io.h"

using namespace mp;

#if defined(MP_COVERAGE)
__device__ unsigned long long mp_cov_bits[mp::COV_WORDS];      // function-hit bitmap of the coverage build (mp_cov.h)
#endif

namespace {

// Kernels other than k_mp_frames: one warp per city, MP_EPC cities (warps) per CTA.  One per CTA: a CTA's
// slot is released the moment its city is done instead of waiting for the slower of two.
constexpr int MP_EPC = 1;
constexpr int MP_THREADS = 32 * MP_EPC;
constexpr int ENV_STRIDE = (EnvLayout::TOTAL + 127) & ~127;

struct MpBatch {
    int n_envs, device;
    uint8_t *blocks;          // [n_envs * ENV_STRIDE]
    uint16_t *anim;           // [1024]
    int32_t *i32_a, *i32_b;   // staging [n_envs * 4]
    uint8_t *mask;            // [n_envs]
    uint8_t *host_blk;        // pinned, one env block
    int32_t *host_i32;        // pinned staging [n_envs * 4]
    char err[256];
    int passes;               // host copy of simPasses (corecontrol.py:129 sets 100)
    int frames_per_launch;    // simFrames per k_mp_frames launch (phase lockstep across the batch)
    int heavy_div;            // the n_envs / heavy_div most expensive envs form the heavy group (0: none)
    int fepc, heavy_fepc;     // envs per CTA of k_mp_frames (frame lockstep inside the CTA), for the heavy group
    unsigned long long *cycles;   // [n_envs] SM cycles each env spent in k_mp_frames since the last read (profiling)
    unsigned int *cost;       // [n_envs] cycles of the current step (schedule key for the next one)
    int *perm;                // [n_envs] env ids, most expensive first
    unsigned long long *cls_sum;   // [SCHED_CLASSES] summed cost of the envs of a schedule class
    unsigned int *cls_cnt;    // [SCHED_CLASSES] how many
    int *perm_pre;            // [n_envs] env ids ordered by the tool of this step's action (k_mp_env_pre of a one-group step)
    int *ncls_dev, *ncls_host;     // number of non-empty schedule classes of the last rebuild (device / pinned host copy)
    int big_groups, big_frames;    // env groups / frames per launch of a single-class batch of >= 32-env CTAs
    int cur_frames;                // frames per launch of the env step being enqueued (0: frames_per_launch)
    int pre_t, pre_l;              // its CTA size and envs per warp
    int pre_simt;                  // the action kernel runs one THREAD per env, four envs to a warp (k_mp_env_pre_t; MPB_PRE_SIMT=0: a warp per env)
    int resched_mid;               // rebuild the order between the two simTicks of a one-group step (MPB_RESCHED_MID=0: off)
    int timing;                    // MPB_TIMING: events at the phase boundaries of a one-group step (mpb_step_phase_times)
    cudaEvent_t ev_ph[6];
    unsigned long long *skey, *skey_out;    // [n_envs] sort keys of the schedule (36 bits used; the tool-order sort of k_mp_env_pre reuses the storage as 32-bit keys)
    int *sval;                // [n_envs] env ids (sort values in)
    void *sort_tmp; size_t sort_tmp_bytes;
    int use_sched;
    int sms;                  // SMs of the device
    int scan_block;           // k_mp_frames: phases after which there is no CTA barrier (bit mask; 0x00fe: the map-scan phases 1..8 free-run inside a CTA)
    unsigned int *prof;       // per-env per-frame cycles [n_envs][256] (profiling, mpb_frame_profile)
    cudaError_t launch_err;   // first error of a launch enqueued by launch_frames since the last check
    // per-episode seeds derived on the device (gym_episode_seed) and the device-side auto-reset of mpb_env_step*
    long long run_seed; int env_offset;
    int auto_reset, auto_terrain;
    // terrain prefetch: the map of every env's NEXT episode, generated ahead of time on a low-priority stream
    uint16_t *next_map;       // [n_envs][12000]
    int32_t *next_ep;         // [n_envs] episode index next_map holds (-1: none)
    unsigned int *pf_stats;   // [2] resets served from a prepared map / generated in line
    cudaStream_t pf_stream; cudaEvent_t ev_pf;
    int *rlist, *rcount;      // [n_envs] envs being reset by the current step, [1] how many
    uint8_t *amask;           // [n_envs] the same as a byte mask
    int32_t *xt_rnd; size_t xt_rnd_bytes;    // device buffer of mpb_env_extinguish's draws (grown on demand)
    double *counters_dev;     // [n_envs * MPB_ENV_NUM_COUNTERS] staging of mpb_env_get_counters
    // env groups on their own streams: while one group's launch drains its slowest envs the other
    // groups' launches keep the SMs busy (the per-env dependency chain is the only real one)
    int groups;
    cudaStream_t gstream[8];
    cudaEvent_t ev_start, ev_done[8], ev_ar0, ev_ar1;     // ev_ar*: around the auto-reset part of the last step
    // gym layer (mpb_env_*)
    int configured, shadow_stride;
    GymCfg cfg;
    uint8_t *shadow;          // [n_envs * shadow_stride]
    float *obs;               // [n_envs, 32, W, H]
    float *obs_out;           // where the next observations go: `obs` or a caller's buffer (mpb_env_set_obs_buffer)
    float *reward;            // [n_envs]
    uint8_t *done;            // [n_envs]
    int64_t *act_dev;         // [n_envs]
    int64_t *h_act;           // pinned
    float *h_reward;          // pinned
    uint8_t *h_done;          // pinned
    uint8_t *host_shadow;     // pinned, one env
};

__device__ __forceinline__ int device_env_index() { return blockIdx.x * MP_EPC + (threadIdx.x >> 5); }

constexpr int SCAL_WORDS = (int)((sizeof(Scalars) + 3) / 4);
constexpr int SCAL_BYTES = SCAL_WORDS * 4;
constexpr int MP_DYN_SMEM = MP_EPC * SCAL_BYTES;    // dynamic shared memory of the one-warp-per-env kernels

// One Env per warp, in shared memory (see Coop), together with the env's scalar block: every kernel works on
// a shared-memory copy of the scalars (ScalPtr tells the compiler so) and writes it back when it is done.
// Lane 0 fills the Env in; every lane then uses it.
// The scalar copies live at the start of the kernel's DYNAMIC shared memory (mp_dyn_smem, declared in mp_state.h):
// ScalPtr addresses them as an offset from that symbol, which is how the compiler knows they are shared memory.
#define MP_SHARED_ENV() __shared__ Env s_env[MP_EPC]; \
    Env &e = s_env[threadIdx.x >> 5]; uint32_t *const sw = reinterpret_cast<uint32_t *>(mp_dyn_smem) + (threadIdx.x >> 5) * SCAL_WORDS

// [ALL lanes] bind the warp's Env to env `env` and stage its scalars in `sw` (shared memory)
__device__ __forceinline__ void bind_env_of(Env &e, uint8_t *blocks, int env, const uint16_t *anim, int *red, uint32_t *sw)
{
    uint8_t *blk = blocks + (size_t)env * ENV_STRIDE;
    const uint32_t *gs = reinterpret_cast<const uint32_t *>(blk + EnvLayout::SCALARS);
    for (int i = threadIdx.x & 31; i < SCAL_WORDS; i += 32) sw[i] = gs[i];
    if ((threadIdx.x & 31) == 0) {
        bind_env(e, blk);
        e.s = reinterpret_cast<Scalars *>(sw);
        e.anim = anim;
        e.c.red = red + (threadIdx.x >> 5) * 8;
    }
    __syncwarp();
}
// [ALL lanes] write the staged scalars back to the env's block
__device__ __forceinline__ void unstage_scalars(Env &e, const uint32_t *sw)
{
    __syncwarp();
    uint32_t *gs = reinterpret_cast<uint32_t *>(e.base() + EnvLayout::SCALARS);
    for (int i = threadIdx.x & 31; i < SCAL_WORDS; i += 32) gs[i] = sw[i];
}
__device__ __forceinline__ void bind_device_env(Env &e, uint8_t *blocks, const uint16_t *anim, int *red, uint32_t *sw)
{
    bind_env_of(e, blocks, device_env_index(), anim, red, sw);
}

#define MP_ENV_GUARD() const int env = device_env_index(); if (env >= n_envs) return; const int lane = threadIdx.x & 31; (void)lane

__global__ void __launch_bounds__(MP_THREADS) k_mp_construct(uint8_t *blocks, const uint16_t *anim, int n_envs)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    env_construct(e, blocks + (size_t)device_env_index() * ENV_STRIDE);
    unstage_scalars(e, sw);
}

// mode 0: clearMap; mode 1: generateSomeCity(terrain_seed) then seedRandom(sim_seed)
__global__ void __launch_bounds__(MP_THREADS)
k_mp_reset(uint8_t *blocks, const uint16_t *anim, int n_envs, const int32_t *terrain_seeds,
           const int32_t *sim_seeds, const uint8_t *mask, int mode)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    if (mask && !mask[env]) return;
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    if (mode == 0) clear_map(e);
    else generate_some_city(e, terrain_seeds[env], sim_seeds[env]);
    unstage_scalars(e, sw);
}

__global__ void __launch_bounds__(MP_THREADS)
k_mp_tool(uint8_t *blocks, const uint16_t *anim, int n_envs, const int32_t *txy, int32_t *results)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    if (lane == 0) {
        const int32_t *a = txy + (size_t)env * 3;
        int r = (a[0] < 0) ? (int)TR_FAILED : tool_down(e, a[0], a[1], a[2]);
        if (results) results[env] = r;
    }
    unstage_scalars(e, sw);
}

// k_mp_env_pre is one warp per env running the clear-then-build logic of ITS action: 14 k instructions of tool code,
// a different tool in every warp, and 55-60 % of its stall samples are no_instruction (profiles/r02_notes.md).  In
// tool order neighbouring CTAs -- the ones an SM holds at the same time -- run the same tool.
__global__ void k_mp_pre_keys(const int64_t *actions, int n, int cells, unsigned int *keys, int *vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a = actions[i];
    keys[i] = a < 0 ? 31u : (unsigned int)min((long long)30, a / cells);
    vals[i] = i;
}

// A group of envs = the slots { off + (w / chunk) * chunk * stride + w % chunk : w < count } of the schedule:
// whole chunks of `chunk` consecutive slots (one k_mp_frames CTA's worth), dealt out to the groups in turn, so that
// the envs of a CTA are neighbours in the schedule -- same class, similar cost.
struct GroupGeom { int off, stride, count, chunk; };
__host__ __device__ __forceinline__ int geom_slot(const GroupGeom &q, int w)
{
    return q.chunk <= 1 ? w * q.stride + q.off : q.off + (w / q.chunk) * q.chunk * q.stride + (w % q.chunk);
}

// `count` simFrames (simLoop(true), main.cpp:403-425) for every env, frames [first, first + count) of
// a simTick of simPasses frames.  All envs of the batch run the same frames in the same launch, so
// every warp on the GPU is in (nearly) the same simulation phase and executes the same code: the
// engine's instruction footprint (~0.4 MB) is far beyond the SM instruction caches, and a launch
// per few dozen frames keeps the live footprint to a few phases' worth (profiles/r01_notes.md).
// animate_before != 0 runs tickEngine's animateTiles first (micropolisgtkengine.py:542-552).
// MP_FEPC = envs (warps) per CTA.  MP_FEPC > 1 adds a CTA barrier per frame: the warps of a CTA run
// the same frame -- the same simulation phase, hence the same code -- at the same time, so an SM's
// instruction caches hold 32 / MP_FEPC phases' worth of code instead of up to 32 (no_instruction was
// 42 % of all stall samples with free-running warps; profiles/r01_notes.md).  The heavy-first order
// puts envs of similar cost next to each other, which keeps the barrier waits short.
#ifndef MP_FWARPS
#define MP_FWARPS 32          // resident warps per SM the frame kernel is compiled for (register budget)
#endif
// MP_PROF: the per-env per-frame cycle profile (mpb_frame_profile) is compiled into its own instantiation; the
// product launches carry neither the store nor its test (1.1 % of the issued instructions, profiles/r02_notes.md).
template <int MP_FEPC, bool MP_PROF>
__global__ void __launch_bounds__(32 * MP_FEPC, MP_FWARPS / MP_FEPC)
k_mp_frames(uint8_t *blocks, const uint16_t *anim, int n_envs, int first, int count, int animate_before,
            int respect_passes, unsigned long long *cycles, unsigned int *cost, const int *perm, GroupGeom q,
            int scan_block, unsigned int *prof, int prof_off, const uint8_t *mask, const int *gcount_dev)
{
    __shared__ int red[8 * MP_FEPC];
    // the scalar block (census counters, cycle counters, RNG state, valves ...) is what the serial
    // lane touches on almost every instruction: it lives in shared memory for the whole launch
    __shared__ Env s_env[MP_FEPC];
    __shared__ int s_ph0;
    // list-driven launches (device-side auto-reset: `perm` is the list of envs being reset and *gcount_dev its
    // length, unknown to the host) use a fixed grid and walk the list; every other launch has one CTA per
    // MP_FEPC slots and runs the loop body once
    const int gcount = gcount_dev ? *gcount_dev : q.count;
    for (int cta = blockIdx.x; cta * MP_FEPC < gcount; cta += gridDim.x) {
    const long long t_begin = clock64();
    // heavy-first: slot k of the grid runs the k-th most expensive env of the previous step, so the
    // long envs (flood / fire fields) start in the first wave instead of forming the launch's tail
    const int widx = cta * MP_FEPC + (threadIdx.x >> 5);
    const int slot = geom_slot(q, widx);
    bool live = widx < gcount && slot < n_envs;
    int env = 0;
    if (live) {
        env = perm ? perm[slot] : slot;
        if (mask && !mask[env]) live = false;             // masked step (mpb_env_step_masked): this env sits it out
    }
    if (MP_FEPC == 1 && !live) continue;
    if (MP_FEPC > 1 && mask && !__syncthreads_or(live)) continue;      // nobody of this CTA takes part
    const int lane = threadIdx.x & 31;
    Env &e = s_env[threadIdx.x >> 5];
    uint32_t *sw = reinterpret_cast<uint32_t *>(mp_dyn_smem) + (threadIdx.x >> 5) * SCAL_WORDS;
    if (live) {
        bind_env_of(e, blocks, env, anim, red, sw);
        if (animate_before && e.s->doAnimation && !e.s->tilesAnimated) animate_tiles(e);
    }
    // Frame lockstep: a CTA barrier after every frame keeps the warps of the CTA in the same simulation
    // phase, i.e. in the same code.  The eight map-scan phases (1..8) run the same code, so they need no
    // lockstep among themselves: with scan_block the barrier is skipped after phases 1..7.  The decision
    // uses the CTA's nominal phase (warp 0's env at the start of the launch), so every warp -- also one
    // whose env was reset to another phase -- executes the same barriers.  (A bounded-skew variant over an
    // mbarrier ring, warps running up to 7 frames ahead, was 3-17 % slower and its polling loop was 59 %
    // of the issued instructions: profiles/r01_notes.md.)
    if (MP_FEPC > 1) {
        if (threadIdx.x == 0) s_ph0 = live ? (e.s->initSimLoad ? 0 : (e.s->phaseCycle & 15)) : 0;
        __syncthreads();
    }
    const int ph0 = MP_FEPC > 1 ? s_ph0 : 0;
    const long long t_loop = clock64();
    unsigned int busy = 0;
    // frames this env runs in this launch, and the gym path's common case kept in registers: at simSpeed 3
    // every frame simulates, so speedCycle and phaseCycle are plain counters (written back after the loop)
    int nrun = 0;
    bool fast = false;
    int phase = 0, sc = 0, psc = 0;
    if (live) {
        nrun = count;
        if (respect_passes) nrun = e.s->simSpeed ? min(count, max(0, (int)e.s->simPasses - first)) : 0;
        fast = e.s->simSpeed == 3 && !e.s->initSimLoad;
        phase = e.s->phaseCycle & 15;
        sc = e.s->speedCycle;
    }
    MP_NOUNROLL for (int i = 0; i < count; i++) {
        const bool go = i < nrun;
        const unsigned int t0 = (unsigned int)clock64();
        if (go) {
            if (fast) {
                sim_phase(e, phase);
                phase = (phase + 1) & 15;
                sc = (sc + 1) & 1023;                                  // simFrame: if (++speedCycle > 1023) speedCycle = 0
                // moveObjects: with an empty sprite list it only counts spriteCycle, which nothing but the sprite
                // handlers reads -- the count stays in a register until a sprite exists (or the launch ends)
                if (lane == 0) {
                    if (e.s->spriteHead < 0) psc++;
                    else {
                        if (psc) { e.s->spriteCycle = (int16_t)(e.s->spriteCycle + psc); psc = 0; }
                        move_objects(e);
                    }
                }
                __syncwarp();
            } else {
                const int ran = sim_loop_head(e);
                sim_loop_tail(e, ran);
            }
        }
        const unsigned int dt = (unsigned int)clock64() - t0;          // this env's own time, barrier waits excluded
        busy += dt;
        if (MP_PROF && go && lane == 0) prof[(size_t)env * 256 + prof_off + first + i] = dt;
        if (MP_FEPC > 1) {
            const int ph = (ph0 + i) & 15;
            if (!((scan_block >> ph) & 1)) __syncthreads();
        }
    }
    if (live) {
        if (fast && lane == 0 && nrun > 0) {
            e.s->phaseCycle = (int16_t)phase;
            e.s->speedCycle = (int16_t)sc;
            e.s->spriteCycle = (int16_t)(e.s->spriteCycle + psc);
        }
        unstage_scalars(e, sw);
        if (lane == 0) {
            busy += (unsigned int)(t_loop - t_begin);                   // state load, animateTiles
            if (cycles) cycles[env] += busy;
            if (cost) cost[env] += busy >> 8;
        }
    }
    if (gcount_dev) { if (MP_FEPC > 1) __syncthreads(); else __syncwarp(); }      // the shared Env / scalars are reused
    }
}

__global__ void __launch_bounds__(MP_THREADS)
k_mp_call(uint8_t *blocks, const uint16_t *anim, int n_envs, int fn, int a, int b)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    env_call(e, fn, a, b);
    unstage_scalars(e, sw);
}

__global__ void k_mp_set_scalar(uint8_t *blocks, int n, int which, const int32_t *vals, int64_t v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Scalars *s = (Scalars *)(blocks + (size_t)i * ENV_STRIDE + EnvLayout::SCALARS);
    switch (which) {
    case 0: s->rng = (uint32_t)vals[i]; break;
    case 1: s->totalFunds = (int)v; break;
    case 2: s->simPasses = (int)v; break;
    case 3: s->enableDisasters = (uint8_t)v; break;
    default: break;
    }
}

// The order in which the envs are dealt out to the CTAs of a step ("schedule").  Envs of one k_mp_frames CTA run in
// frame lockstep, which pays off only when they execute the same code in the same frame:
//  * the same phase of the 16-phase cycle: an env that was reset (or stepped under a mask) on another step is half a
//    cycle away from its neighbours (200 frames per env-step = 8 mod 16) -- its map-scan block would meet the others'
//    census / power-scan frames (measured: 38.8 ms per step instead of 20 with episodes of 200 steps);
//  * the same simulation cycle count: the periodic scans (power every 5 cycles, pollution / land value every 17,
//    crime 18, population density 19, fire coverage 20, the census every 4 and tax / evaluation every 48 city-time
//    units) fall into the same frames only for envs with the same cityTime; in a CTA of mixed ages some env runs
//    one of them in almost every cycle and the CTA pays for it every time (28.6 ms instead of 20.4, no reset in
//    sight: tools/episode_probe2.py).
// So the primary sort key is the env's schedule CLASS = hash(cityTime, phase) -- envs reset on the same step share
// it --, classes ordered by their mean cost (most expensive first: long CTAs start early in a launch), and inside a
// class the envs by their own cost of the previous step, most expensive first.  With one class (all envs reset
// together: the plain rollout) this is the heavy-first order by cost alone.
constexpr int SCHED_CLASS_BITS = 12;
constexpr int SCHED_CLASSES = 1 << SCHED_CLASS_BITS;
__device__ __forceinline__ int env_phase(const uint8_t *blocks, int env)
{
    const Scalars *sc = reinterpret_cast<const Scalars *>(blocks + (size_t)env * ENV_STRIDE + EnvLayout::SCALARS);
    return sc->initSimLoad ? 0 : (sc->phaseCycle & 15);
}
__device__ __forceinline__ unsigned int env_sched_class(const uint8_t *blocks, int env)
{
    const Scalars *sc = reinterpret_cast<const Scalars *>(blocks + (size_t)env * ENV_STRIDE + EnvLayout::SCALARS);
    const unsigned int phase = sc->initSimLoad ? 0u : (unsigned int)(sc->phaseCycle & 15);
    const unsigned int k = ((unsigned int)sc->cityTime * 16u + phase) * 2654435761u;
    return k >> (32 - SCHED_CLASS_BITS);
}
__global__ void k_mp_sched_identity(int *perm, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = i;
}
__global__ void k_mp_sched_class(const unsigned int *cost, const uint8_t *blocks, int n, unsigned long long *cls_sum,
                                 unsigned int *cls_cnt, int *ncls)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int c = env_sched_class(blocks, i);
    atomicAdd(&cls_sum[c], (unsigned long long)cost[i]);
    if (atomicAdd(&cls_cnt[c], 1u) == 0u) atomicAdd(ncls, 1);
}
// key = [phase: 4 bits][mean cost of the class, inverted: 10 bits][class: 12 bits][own cost, inverted: 10 bits]; cost =
// cycles >> 8 of the previous step, bucketed by 2^5 (2^13 busy cycles per bucket; the mean env is at ~300).  The phase
// comes first so that the classes that meet in a CTA at a class boundary (a 164-env class is 5.1 CTAs) are at least in
// the same phase: such a CTA then pays the periodic scans of two calendars, not the maximum of two different phases
// in every frame.
__global__ void k_mp_sched_keys(unsigned int *cost, const uint8_t *blocks, int n, const unsigned long long *cls_sum,
                                const unsigned int *cls_cnt, unsigned long long *keys, int *vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int c = env_sched_class(blocks, i);
    const unsigned int mean = (unsigned int)(cls_sum[c] / (unsigned long long)max(1u, cls_cnt[c]));
    const unsigned int q = min(1023u, mean >> 5), own = min(1023u, cost[i] >> 5);
    const unsigned int lo = ((1023u - q) << 22) | (c << 10) | (1023u - own);
    keys[i] = ((unsigned long long)env_phase(blocks, i) << 32) | lo;
    vals[i] = i;
    cost[i] = 0;
}

__device__ __forceinline__ void bind_device_gym(Gym &gy, uint8_t *shadow, int stride, const GymCfg *cfg)
{
    bind_gym(gy, shadow + (size_t)device_env_index() * stride, cfg);
}

// ---- terrain prefetch ------------------------------------------------------------------------------------------
// generateMap is 2.5 ms of single-warp latency, and in an auto-reset it sits on the step's critical path with the GPU
// nearly idle (a few hundred envs reset per step).  The seeds of an env's next episode are known the moment the current
// one starts, so its map is generated ahead of time, into next_map, by a low-priority kernel that runs in the gaps of
// the following steps; the reset then copies 24 KB instead of generating.  next_ep[env] says which episode the
// prepared map belongs to and is written last; a reset that finds another value generates in line.
__device__ __forceinline__ const uint16_t *prepared_terrain(const uint16_t *next_map, const int32_t *next_ep,
                                                            unsigned int *pf_stats, int env, int ep)
{
    if (!next_map) return nullptr;
    int have = 0;
    if ((threadIdx.x & 31) == 0) {
        have = *reinterpret_cast<const volatile int32_t *>(next_ep + env) == ep;
        atomicAdd(&pf_stats[have ? 0 : 1], 1u);
    }
    have = __shfl_sync(0xffffffffu, have, 0);
    __threadfence();
    return have ? next_map + (size_t)env * NTILES : nullptr;
}
__global__ void __launch_bounds__(MP_THREADS)
k_mp_prefetch_terrain(const uint8_t *shadow, int stride, GymCfg cfg, int n, uint16_t *next_map, int32_t *next_ep,
                      long long run_seed, int env_offset)
{
    __shared__ int red[8 * MP_EPC];
    MP_SHARED_ENV();
    const int lane = threadIdx.x & 31;
    for (int env = blockIdx.x; env < n; env += gridDim.x) {
        int ep = 0, need = 0;
        if (lane == 0) {
            Gym gy;
            bind_gym(gy, const_cast<uint8_t *>(shadow) + (size_t)env * stride, &cfg);
            ep = *reinterpret_cast<const volatile int32_t *>(&gy.g->episode);
            need = *reinterpret_cast<const volatile int32_t *>(next_ep + env) != ep;
        }
        ep = __shfl_sync(0xffffffffu, ep, 0);
        need = __shfl_sync(0xffffffffu, need, 0);
        if (!need) continue;
        for (int i = lane; i < SCAL_WORDS; i += 32) sw[i] = 0;       // a private scalar block: the generator only uses its RNG
        if (lane == 0) {
            bind_env(e, reinterpret_cast<uint8_t *>(next_map + (size_t)env * NTILES));     // map() is the block's first field
            e.s = reinterpret_cast<Scalars *>(sw);
            e.anim = nullptr;
            e.c.red = red;
        }
        __syncwarp();
#if MP_WARP_COOP                                             // (the host pass of nvcc does not see the warp-cooperative generator)
        generate_map_coop(e, gym_episode_seed((uint64_t)run_seed, (uint64_t)(env_offset + env), (uint64_t)ep, 0));
#endif
        __threadfence();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile int32_t *>(next_ep + env) = ep;
        __syncwarp();
    }
}
static_assert(EnvLayout::MAP == 0, "the prefetch kernel binds an Env whose block is just the map");

// terrain_seeds == NULL with terrain != 0: the seeds of the env's next episode are derived here (gym_episode_seed)
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_reset_begin(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg,
                     const int32_t *terrain_seeds, const int32_t *sim_seeds, const uint8_t *mask, int terrain,
                     long long run_seed, int env_offset, const uint16_t *next_map, const int32_t *next_ep, unsigned int *pf_stats)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    if (mask && !mask[env]) return;
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    Gym gy;
    bind_device_gym(gy, shadow, stride, &cfg);
    int ts = 0, ss = 0;
    const uint16_t *prepared = nullptr;
    if (terrain && terrain_seeds) { ts = terrain_seeds[env]; ss = sim_seeds[env]; }
    else if (terrain) {
        const int ep = gy.g->episode;
        ts = gym_episode_seed((uint64_t)run_seed, (uint64_t)(env_offset + env), (uint64_t)ep, 0);
        ss = gym_episode_seed((uint64_t)run_seed, (uint64_t)(env_offset + env), (uint64_t)ep, 1);
        prepared = prepared_terrain(next_map, next_ep, pf_stats, env, ep);
    }
    __syncwarp();
    gym_reset_begin(e, gy, terrain, ts, ss, prepared);
    if (lane == 0 && terrain && !terrain_seeds) gy.g->episode++;
    unstage_scalars(e, sw);
}

// ---- device-side auto-reset (subproc_vec_env.py:16-21: `if done: ob = env.reset()` inside the worker) ----------
// The envs whose episode ended in the step that just ran, as a list + count + byte mask; nothing goes to the host.
__global__ void k_mp_env_auto_list(const uint8_t *done, const uint8_t *step_mask, int n, int *list, int *count, uint8_t *amask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool r = done[i] != 0 && (!step_mask || step_mask[i] != 0);
    amask[i] = r ? 1 : 0;
    if (r) list[atomicAdd(count, 1)] = i;
}
// reset, first half, for the envs of the list (a fixed grid walks it: the host does not know its length)
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_reset_list(uint8_t *blocks, const uint16_t *anim, uint8_t *shadow, int stride, GymCfg cfg, const int *list,
                    const int *count, int terrain, long long run_seed, int env_offset, const uint16_t *next_map,
                    const int32_t *next_ep, unsigned int *pf_stats)
{
    __shared__ int red[8 * MP_EPC];
    MP_SHARED_ENV();
    const int lane = threadIdx.x & 31;
    const int n = *count;
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int env = list[k];
        bind_env_of(e, blocks, env, anim, red, sw);
        Gym gy;
        bind_gym(gy, shadow + (size_t)env * stride, &cfg);
        const int ep = gy.g->episode;
        const uint16_t *prepared = terrain ? prepared_terrain(next_map, next_ep, pf_stats, env, ep) : nullptr;
        __syncwarp();
        gym_reset_begin(e, gy, terrain,
                        terrain ? gym_episode_seed((uint64_t)run_seed, (uint64_t)(env_offset + env), (uint64_t)ep, 0) : 0,
                        terrain ? gym_episode_seed((uint64_t)run_seed, (uint64_t)(env_offset + env), (uint64_t)ep, 1) : 0,
                        prepared);
        if (lane == 0 && terrain) gy.g->episode = ep + 1;
        unstage_scalars(e, sw);
        __syncwarp();
    }
}
// reset, second half after the tick, for the envs of the list: metrics, funds, observation; reward / done stay
// those of the step that ended the episode
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_post_list(uint8_t *blocks, const uint16_t *anim, uint8_t *shadow, int stride, GymCfg cfg, float *obs,
                   float *reward, uint8_t *done, const int *list, const int *count)
{
    __shared__ int red[8 * MP_EPC];
    MP_SHARED_ENV();
    const int n = *count;
    const size_t on = (size_t)gym_obs_channels(cfg) * cfg.W * cfg.H;
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int env = list[k];
        bind_env_of(e, blocks, env, anim, red, sw);
        Gym gy;
        bind_gym(gy, shadow + (size_t)env * stride, &cfg);
        gym_reset_post(e, gy, obs + env * on, reward + env, done + env, true);
        unstage_scalars(e, sw);
        __syncwarp();
    }
}

// MicropolisControl.updateMap (corecontrol.py:193-203): every window cell of the shadow map re-read from the engine map
// (after a city was loaded into the engine: mpb_set_field MPF_MAP / cty.load_env)
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_update_map(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg, const uint8_t *mask)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    if (mask && !mask[env]) return;
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    Gym gy;
    bind_device_gym(gy, shadow, stride, &cfg);
    gym_update_map(e, gy);
    unstage_scalars(e, sw);
}

// TileMap.addZoneBot through MicropolisControl.doBotTool, no tick (powerPuzzle / randomStaticStart set-up)
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_action(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg,
                const int64_t *actions, int static_build, int full, int32_t *ok)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    Gym gy;
    bind_device_gym(gy, shadow, stride, &cfg);
    if (lane == 0) {
        int r = gym_apply_action(e, gy, actions[env], static_build != 0, full != 0);
        if (ok) ok[env] = r;
    }
    unstage_scalars(e, sw);
}

// MicropolisPaintEnv.step, part 1 (paintenv.py:43-52): one tool per window cell, then setFunds
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_pre_paint(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg,
                   const uint8_t *actions, const int *perm, GroupGeom q)
{
    __shared__ int red[8 * MP_EPC];
    if (device_env_index() >= q.count) return;
    const int slot = geom_slot(q, device_env_index());
    if (slot >= n_envs) return;
    const int env = perm ? perm[slot] : slot;
    const int lane = threadIdx.x & 31;
    MP_SHARED_ENV();
    bind_env_of(e, blocks, env, anim, red, sw);
    Gym gy;
    bind_gym(gy, shadow + (size_t)env * stride, &cfg);
    if (lane == 0) {
        gym_paint_action(e, gy, actions + (size_t)env * cfg.W * cfg.H);
        e.s->totalFunds = cfg.init_funds;
    }
    unstage_scalars(e, sw);
}

// Extinguisher wrapper (wrappers.py:10-160): init_age_array / one extinction event per env
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_track_ages(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    Gym gy;
    bind_device_gym(gy, shadow, stride, &cfg);
    gym_init_age_array(e, gy);
}
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_extinguish(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg, int type,
                    int n_dels, const int32_t *rnd, const uint8_t *mask, int32_t *dels_out, double interval)
{
    __shared__ int red[8 * MP_EPC];
    MP_ENV_GUARD();
    if (mask && !mask[env]) { if (lane == 0) dels_out[env] = 0; return; }
    Gym gy;
    bind_device_gym(gy, shadow, stride, &cfg);
    if (interval != 0.0) {
        // Extinguisher.step (wrappers.py:35-40): `if self.num_step % self.extinction_interval == 0` on the env's OWN step
        // count.  An env that was just reset (num_step == 0) is skipped: in the reference the wrapper sits below the
        // worker's auto-reset, so the event would have hit the city the reset then replaced.
        const int k = gy.g->num_step;
        if (k <= 0 || fmod((double)k, interval) != 0.0) { if (lane == 0) dels_out[env] = 0; return; }
    }
    MP_SHARED_ENV();
    bind_device_env(e, blocks, anim, red, sw);
    if (lane == 0) dels_out[env] = gym_extinguish(e, gy, type, n_dels, rnd + (size_t)env * (n_dels + 2));
    unstage_scalars(e, sw);
}

// MicropolisEnv.step, part 1 (env.py:448-466): decoded action through TileMap.addZoneBot, setFunds
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_pre(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg,
             const int64_t *actions, int static_build, const int *perm, GroupGeom q, const uint8_t *mask)
{
    __shared__ int red[8 * MP_EPC];
    if (device_env_index() >= q.count) return;
    const int slot = geom_slot(q, device_env_index());
    if (slot >= n_envs) return;
    const int env = perm ? perm[slot] : slot;
    if (mask && !mask[env]) return;
    const int lane = threadIdx.x & 31;
    MP_SHARED_ENV();
    bind_env_of(e, blocks, env, anim, red, sw);
    __shared__ __align__(16) uint16_t s_ovl[MP_EPC][272];      // the tool overlay (256 values + mask) on chip
    if (lane == 0) e.ovl = s_ovl[threadIdx.x >> 5];
    __syncwarp();
    Gym gy;
    bind_gym(gy, shadow + (size_t)env * stride, &cfg);
    if (lane == 0) gym_step_pre(e, gy, actions[env], static_build != 0);
    unstage_scalars(e, sw);
}

// The same step part with one THREAD per env, a few envs to a warp: gym_step_pre and everything below it (TileMap
// bookkeeping, toolDown, connect / fixZone, the tool overlay) is single-lane code without a warp collective in it, so
// the first lanes of a warp can run it for an env each (in tool order: the same tool).  The lanes share next to
// nothing -- 16 or 32 envs per warp take 16 or 32 times as long as one, also when ordered by (tool, zone at the
// target) -- but a warp that interleaves a few diverged instruction streams hides their latencies from one another:
// 1.91 -> 1.37 ms for 32 768 envs (128-thread CTAs: 2 envs per warp 1.71, 4 envs 1.51, 8 envs 1.56, 16 envs 2.89 ms;
// profiles/r02_notes.md).  Scalars: staged per thread in dynamic shared memory, copied in and out by the whole warp
// (coalesced); the tool overlay stays in the env block's scratch.
constexpr int PRE_T = 64;                       // threads per CTA (MPB_PRE_T)
constexpr int PRE_L = 6;                        // lanes of every warp that take an env each (MPB_PRE_L; 64 x 6: 1.37 ms, 128 x 4: 1.47 ms)
__global__ void __launch_bounds__(256)
k_mp_env_pre_t(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg,
               const int64_t *actions, const uint8_t *paint, int static_build, const int *perm, GroupGeom q,
               const uint8_t *mask, int PRE_L)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = (blockIdx.x * (blockDim.x / 32) + warp) * PRE_L + lane;
    uint32_t *wsw = reinterpret_cast<uint32_t *>(mp_dyn_smem) + (size_t)warp * PRE_L * SCAL_WORDS;
    int env = -1;
    if (lane < PRE_L && idx < q.count) {
        const int slot = geom_slot(q, idx);
        if (slot < n_envs) {
            env = perm ? perm[slot] : slot;
            if (mask && !mask[env]) env = -1;
        }
    }
    MP_NOUNROLL for (int k = 0; k < PRE_L; k++) {
        const int ek = __shfl_sync(0xffffffffu, env, k);
        if (ek < 0) continue;
        const uint32_t *gs = reinterpret_cast<const uint32_t *>(blocks + (size_t)ek * ENV_STRIDE + EnvLayout::SCALARS);
        for (int i = lane; i < SCAL_WORDS; i += 32) wsw[k * SCAL_WORDS + i] = gs[i];
    }
    __syncwarp();
    if (env >= 0) {
        Env e;
        bind_env(e, blocks + (size_t)env * ENV_STRIDE);
        e.s = reinterpret_cast<Scalars *>(wsw + lane * SCAL_WORDS);
        e.anim = anim;
        e.c.red = nullptr;
        Gym gy;
        bind_gym(gy, shadow + (size_t)env * stride, &cfg);
        if (paint) {                                    // MicropolisPaintEnv.step, part 1 (paintenv.py:43-52)
            gym_paint_action(e, gy, paint + (size_t)env * cfg.W * cfg.H);
            e.s->totalFunds = cfg.init_funds;
        } else
            gym_step_pre(e, gy, actions[env], static_build != 0);
    }
    __syncwarp();
    MP_NOUNROLL for (int k = 0; k < PRE_L; k++) {
        const int ek = __shfl_sync(0xffffffffu, env, k);
        if (ek < 0) continue;
        uint32_t *gs = reinterpret_cast<uint32_t *>(blocks + (size_t)ek * ENV_STRIDE + EnvLayout::SCALARS);
        for (int i = lane; i < SCAL_WORDS; i += 32) gs[i] = wsw[k * SCAL_WORDS + i];
    }
}

// MicropolisEnv.step, part 3 (env.py:468-540): observation, metrics, reward, done
__global__ void __launch_bounds__(MP_THREADS)
k_mp_env_post(uint8_t *blocks, const uint16_t *anim, int n_envs, uint8_t *shadow, int stride, GymCfg cfg,
              float *obs, float *reward, uint8_t *done, int is_reset, const uint8_t *mask, const int *perm,
              GroupGeom q)
{
    __shared__ int red[8 * MP_EPC];
    if (device_env_index() >= q.count) return;
    const int slot = geom_slot(q, device_env_index());
    if (slot >= n_envs) return;
    const int env = perm ? perm[slot] : slot;
    const int lane = threadIdx.x & 31;
    (void)lane;
    if (mask && !mask[env]) return;
    MP_SHARED_ENV();
    bind_env_of(e, blocks, env, anim, red, sw);
    Gym gy;
    bind_gym(gy, shadow + (size_t)env * stride, &cfg);
    const size_t on = (size_t)gym_obs_channels(cfg) * cfg.W * cfg.H;
    if (is_reset) gym_reset_post(e, gy, obs + env * on, reward + env, done + env);
    else gym_step_post(e, gy, obs + env * on, reward + env, done + env);
    unstage_scalars(e, sw);
}

__global__ void k_mp_env_set_num_step(uint8_t *shadow, int stride, GymCfg cfg, int n, const int32_t *vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Gym gy;
    bind_gym(gy, shadow + (size_t)i * stride, &cfg);
    gy.g->num_step = vals[i];
}

__global__ void k_mp_env_counters(const uint8_t *blocks, const uint8_t *shadow, int stride, GymCfg cfg, int n, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Scalars *sc = (const Scalars *)(blocks + (size_t)i * ENV_STRIDE + EnvLayout::SCALARS);
    Gym gy;
    bind_gym(gy, const_cast<uint8_t *>(shadow) + (size_t)i * stride, &cfg);
    const GymScalars &g = *gy.g;
    double *o = out + (size_t)i * MPB_ENV_NUM_COUNTERS;
    o[0] = g.num_plants; o[1] = g.num_roads; o[2] = g.n_struct_tiles; o[3] = g.total_traffic;
    o[4] = g.num_step; o[5] = g.tilemap_error; o[6] = g.reward; o[7] = g.done; o[8] = g.curr_funds;
    for (int k = 0; k < GYM_NUM_METRICS; k++) o[9 + k] = g.metrics[k];
    o[15] = sc->spriteOverflow;
    o[16] = g.episode;
}

#define MP_GRID(b) (((b)->n_envs + MP_EPC - 1) / MP_EPC)

// Rebuild the schedule (see env_sched_class) from the phase / city time every env is at now and the cycles the
// previous step cost it, then clear the costs for this step.  The sort itself is cub::DeviceRadixSort (n 32-bit
// keys, ~40 us for 32 768 envs, 0.2 % of a step).
int rebuild_schedule(MpBatch *b, cudaStream_t st)
{
    if (!b->use_sched) return 0;
    const int n = b->n_envs;
    cudaMemsetAsync(b->cls_sum, 0, SCHED_CLASSES * sizeof(unsigned long long), st);
    cudaMemsetAsync(b->cls_cnt, 0, SCHED_CLASSES * sizeof(unsigned int), st);
    cudaMemsetAsync(b->ncls_dev, 0, sizeof(int), st);
    k_mp_sched_class<<<(n + 255) / 256, 256, 0, st>>>(b->cost, b->blocks, n, b->cls_sum, b->cls_cnt, b->ncls_dev);
    // how many classes the batch has, for the host's choice of the group geometry of a LATER step (no sync: the
    // pinned word is read whenever it happens to have arrived)
    cudaMemcpyAsync(b->ncls_host, b->ncls_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
    k_mp_sched_keys<<<(n + 255) / 256, 256, 0, st>>>(b->cost, b->blocks, n, b->cls_sum, b->cls_cnt, b->skey, b->sval);
    size_t bytes = b->sort_tmp_bytes;
    const cudaError_t e = cub::DeviceRadixSort::SortPairs(b->sort_tmp, bytes, b->skey, b->skey_out, b->sval, b->perm, n, 0, 36, st);
    if (e != cudaSuccess && b->launch_err == cudaSuccess) b->launch_err = e;
    return 0;
}


// k_mp_frames with `epc` envs per CTA (1, 8, 16 or 32); prof_off = frame
// offset of this simTick in the per-frame profile rows (0 for the first simTick of an env-step, 100 for the
// second).  `mask`: device mask of the envs that take part (NULL = all).  A launch error is recorded in the handle
// (b->launch_err) and reported by the entry point that enqueued it.
void launch_frames(MpBatch *b, int epc, cudaStream_t st, int first, int count, int animate_before, int respect_passes,
                   const int *perm, GroupGeom q, const uint8_t *mask, int prof_off = 0, bool heavy = false)
{
    const int gcount = q.count;
    (void)heavy;
    if (gcount <= 0) return;
#define MP_FRAMES_ARGS b->blocks, b->anim, b->n_envs, first, count, animate_before, respect_passes, b->cycles, b->cost, perm, \
                       q, b->scan_block, b->prof, prof_off, mask, (const int *)NULL
#define MP_FRAMES_LAUNCH(PROF) do { \
        if (epc >= 32) k_mp_frames<32, PROF><<<(gcount + 31) / 32, 1024, 32 * SCAL_BYTES, st>>>(MP_FRAMES_ARGS); \
        else if (epc >= 16) k_mp_frames<16, PROF><<<(gcount + 15) / 16, 512, 16 * SCAL_BYTES, st>>>(MP_FRAMES_ARGS); \
        else if (epc >= 8) k_mp_frames<8, PROF><<<(gcount + 7) / 8, 256, 8 * SCAL_BYTES, st>>>(MP_FRAMES_ARGS); \
        else k_mp_frames<1, PROF><<<gcount, 32, SCAL_BYTES, st>>>(MP_FRAMES_ARGS); } while (0)
    if (b->prof) MP_FRAMES_LAUNCH(true); else MP_FRAMES_LAUNCH(false);
#undef MP_FRAMES_LAUNCH
#undef MP_FRAMES_ARGS
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess && b->launch_err == cudaSuccess) b->launch_err = le;
}

// C++ simTick (main.cpp:435-443) for the batch: simPasses frames in launches of frames_per_launch.
// Env groups (each on its own stream).  Group 0 holds the `heavy` most expensive envs of the last
// step (flood / fire fields, 5-10x the mean): isolated there they stretch only their own launches,
// while the other groups -- the remaining envs dealt out round robin in cost order -- have tight
// cost distributions and short launch tails.  A group is (offset, stride, count) into the
// heavy-first permutation.
GroupGeom group_geom(const MpBatch *b, int G, int g)
{
    GroupGeom q;
    q.chunk = 1;
    if (G <= 1) { q.off = 0; q.stride = 1; q.count = b->n_envs; return q; }
    int heavy = b->use_sched && b->heavy_div > 0 ? b->n_envs / b->heavy_div : 0;   // ~1.5 % of the batch
    if (heavy < 1 || G < 3) heavy = 0;
    if (heavy > 0 && g == 0) { q.off = 0; q.stride = 1; q.count = heavy; return q; }
    // the other groups: chunks of one CTA's worth of consecutive slots, dealt out in turn
    const int Gn = heavy > 0 ? G - 1 : G, gi = heavy > 0 ? g - 1 : g;
    const int CH = b->fepc > 1 ? b->fepc : 1;
    const int T = b->n_envs - heavy, nchunks = (T + CH - 1) / CH;
    const int mine = nchunks > gi ? (nchunks - gi + Gn - 1) / Gn : 0;
    q.chunk = CH; q.stride = Gn; q.off = heavy + gi * CH; q.count = mine * CH;       // slots >= n_envs are skipped in the kernels
    return q;
}
#define MP_GRID_N(count) (((count) + MP_EPC - 1) / MP_EPC)
// env groups of an env-step (their launch tails overlap on their own streams) and frames per launch: see mpb_create
bool single_class_big(const MpBatch *b) { return b->fepc >= 32 && *reinterpret_cast<volatile int *>(b->ncls_host) <= 1; }
int step_groups(const MpBatch *b) { return single_class_big(b) ? b->big_groups : b->groups; }
int step_frames(const MpBatch *b) { return single_class_big(b) ? b->big_frames : b->frames_per_launch; }

int launch_sim_tick(MpBatch *b, cudaStream_t st, int animate_before, const uint8_t *mask, int G = 1, int g = 0)
{
    const int P = b->passes;
    const int *perm = b->use_sched ? b->perm : NULL;
    const GroupGeom q = group_geom(b, G, g);
    // the heavy group is not kept in phase lockstep: with few, very uneven envs every launch would
    // wait for ITS slowest member (a different env each time), so it runs a whole simTick per launch
    const bool heavy_group = (G > 1 && g == 0 && q.stride == 1 && q.chunk == 1);
    const int F = heavy_group ? (P > 0 ? P : 1) : (b->cur_frames > 0 ? b->cur_frames : b->frames_per_launch);
    const int epc = heavy_group ? b->heavy_fepc : b->fepc;
    if (q.count <= 0) return 0;
    if (P <= 0) {
        if (animate_before)
            launch_frames(b, epc, st, 0, 0, 1, 1, perm, q, mask, 0, heavy_group);
        return 0;
    }
    for (int f = 0; f < P; f += F)
        launch_frames(b, epc, st, f, (P - f < F) ? P - f : F, animate_before && f == 0, 1, perm, q,
                      mask, animate_before ? P : 0, heavy_group);
    return 0;
}
// MicropolisControl.simTick (corecontrol.py:148-152): tickEngine (simTick + animateTiles) + simTick
int rebuild_schedule(MpBatch *b, cudaStream_t st);
int launch_control_sim_tick(MpBatch *b, cudaStream_t st, const uint8_t *mask, int G = 1, int g = 0, bool resched = false)
{
    launch_sim_tick(b, st, 0, mask, G, g);
    // one-group steps: the order is rebuilt between the two simTicks from what the first one cost -- a city that started
    // to flood or burn during it (40 % of all barrier waiting is for such a city) joins the CTAs of its new peers half a
    // step earlier: +4 % (1.77 -> 1.85 M env-ticks/s).  A new order before every 50- or 25-frame launch is worse again
    // (1.81 / 1.70 M: more launch tails, noisier keys), and so is mixing in older costs (profiles/r02_notes.md).
    if (resched) rebuild_schedule(b, st);
    launch_sim_tick(b, st, 1, mask, G, g);
    return 0;
}

// The envs whose `done` flag the step just set are reset before the step returns, all on the device: list of the
// finished envs, reset (seeds of each env's next episode derived in the kernel), MicropolisControl.simTick and the
// reset observation for the envs of the list.  The list kernels run on a fixed grid and walk the list, so a step
// in which no episode ended pays six empty launches (~20 us) and nothing crosses PCIe.
void launch_auto_reset(MpBatch *b, const uint8_t *step_mask, cudaStream_t st)
{
    const int n = b->n_envs;
    int grid = b->sms * 32;
    if (grid > n) grid = n;
    cudaEventRecord(b->ev_ar0, st);
    cudaMemsetAsync(b->rcount, 0, 4, st);
    k_mp_env_auto_list<<<(n + 255) / 256, 256, 0, st>>>(b->done, step_mask, n, b->rlist, b->rcount, b->amask);
    k_mp_env_reset_list<<<grid, MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->shadow, b->shadow_stride, b->cfg, b->rlist,
                                                               b->rcount, b->auto_terrain, b->run_seed, b->env_offset,
                                                               b->next_map, b->next_ep, b->pf_stats);
    const int P = b->passes;
    for (int t = 0; t < 2; t++)                               // MicropolisControl.simTick: simTick, animateTiles, simTick
        k_mp_frames<1, false><<<grid, 32, SCAL_BYTES, st>>>(b->blocks, b->anim, n, 0, P > 0 ? P : 0, t, 1, b->cycles, b->cost, b->rlist,
                                                    GroupGeom{0, 1, 0, 1}, b->scan_block, NULL, 0, NULL, b->rcount);
    k_mp_env_post_list<<<grid, MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->shadow, b->shadow_stride, b->cfg, b->obs_out,
                                                              b->reward, b->done, b->rlist, b->rcount);
    cudaEventRecord(b->ev_ar1, st);
    if (b->next_map && b->auto_terrain) {                    // maps of the next episodes, in the gaps of the following steps
        cudaEventRecord(b->ev_pf, st);
        cudaStreamWaitEvent(b->pf_stream, b->ev_pf, 0);
        k_mp_prefetch_terrain<<<grid, MP_THREADS, MP_DYN_SMEM, b->pf_stream>>>(b->shadow, b->shadow_stride, b->cfg, n, b->next_map,
                                                                              b->next_ep, b->run_seed, b->env_offset);
    }
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess && b->launch_err == cudaSuccess) b->launch_err = le;
}

int fail(MpBatch *b, const char *what, cudaError_t e)
{
    snprintf(b->err, sizeof(b->err), "%s: %s", what, cudaGetErrorString(e));
    return -1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(b, #call, e_); } while (0)
// kernel launches enqueued through launch_frames record their error in the handle
#define CK_LAUNCHES() do { if (b->launch_err != cudaSuccess) { cudaError_t e_ = b->launch_err; b->launch_err = cudaSuccess; return fail(b, "kernel launch", e_); } CK(cudaGetLastError()); } while (0)

int fetch_env(MpBatch *b, int env, Env &e)
{
    if (env < 0 || env >= b->n_envs) { snprintf(b->err, sizeof(b->err), "env %d out of range", env); return -1; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(b->host_blk, b->blocks + (size_t)env * ENV_STRIDE, EnvLayout::TOTAL, cudaMemcpyDeviceToHost));
    bind_env(e, b->host_blk);
    return 0;
}
int store_env(MpBatch *b, int env)
{
    CK(cudaMemcpy(b->blocks + (size_t)env * ENV_STRIDE, b->host_blk, EnvLayout::TOTAL, cudaMemcpyHostToDevice));
    return 0;
}

} // namespace

extern "C" {

mpb_handle mpb_create(int n_envs, int device)
{
    if (n_envs <= 0) return NULL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return NULL;
    cudaSetDevice(device);
    MpBatch *b = (MpBatch *)calloc(1, sizeof(MpBatch));
    b->n_envs = n_envs; b->device = device;
    b->passes = 100;
    b->frames_per_launch = 50;
    if (const char *fp = getenv("MPB_FRAMES_PER_LAUNCH")) { int v = atoi(fp); if (v > 0) b->frames_per_launch = v; }
    b->heavy_div = 64;
    if (const char *fp = getenv("MPB_HEAVY_DIV")) b->heavy_div = atoi(fp);
    // envs per CTA of the frame kernel: 32 (one CTA per SM in lockstep) needs enough CTAs to keep every SM
    // busy for several waves; smaller batches take 16 or 8 (measured: 8 192 envs +5 % with 16, 2 048 +11 % with 8)
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        b->sms = sms;
        b->fepc = n_envs >= sms * 32 * 4 ? 32 : (n_envs >= sms * 16 * 2 ? 16 : 8);
        b->heavy_fepc = b->fepc;
    }
    if (const char *fp = getenv("MPB_FEPC")) b->fepc = atoi(fp);
    if (const char *fp = getenv("MPB_HEAVY_FEPC")) b->heavy_fepc = atoi(fp);
    if (const char *fp = getenv("MPB_CARVEOUT")) {           // experiment: shared-memory share of the L1 (percent)
        cudaFuncSetAttribute(k_mp_frames<16, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(fp));
        cudaFuncSetAttribute(k_mp_frames<8, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(fp));
    }
    b->scan_block = 0x00fe;                         // bit p set: no CTA barrier after a frame of phase p (phases 1..7)
    if (const char *fp = getenv("MPB_SCAN_BLOCK")) b->scan_block = atoi(fp) != 0 ? 0x00fe : 0;
    if (const char *fp = getenv("MPB_NOBAR_MASK")) b->scan_block = (int)strtol(fp, NULL, 0) & 0xffff;
    uint16_t anim[1024];
    for (int i = 0; i < 1024; i++) anim[i] = (uint16_t)i;
    for (int i = 0; i < kAnimPairCount; i++) anim[kAnimPairs[i][0]] = kAnimPairs[i][1];
    bool ok = cudaMalloc(&b->blocks, (size_t)n_envs * ENV_STRIDE) == cudaSuccess
        && cudaMalloc(&b->anim, sizeof(anim)) == cudaSuccess
        && cudaMalloc(&b->i32_a, (size_t)n_envs * 16) == cudaSuccess
        && cudaMalloc(&b->i32_b, (size_t)n_envs * 16) == cudaSuccess
        && cudaMalloc(&b->mask, n_envs) == cudaSuccess
        && cudaMalloc(&b->cycles, (size_t)n_envs * 8) == cudaSuccess
        && cudaMalloc(&b->cost, (size_t)n_envs * 4) == cudaSuccess
        && cudaMalloc(&b->perm, (size_t)n_envs * 4) == cudaSuccess
        && cudaMalloc(&b->cls_sum, SCHED_CLASSES * sizeof(unsigned long long)) == cudaSuccess
        && cudaMalloc(&b->cls_cnt, SCHED_CLASSES * sizeof(unsigned int)) == cudaSuccess
        && cudaMalloc(&b->perm_pre, (size_t)n_envs * 4) == cudaSuccess

        && cudaMalloc(&b->ncls_dev, sizeof(int)) == cudaSuccess
        && cudaMallocHost(&b->ncls_host, sizeof(int)) == cudaSuccess
        && cudaMalloc(&b->skey, (size_t)n_envs * 8) == cudaSuccess
        && cudaMalloc(&b->skey_out, (size_t)n_envs * 8) == cudaSuccess
        && cudaMalloc(&b->sval, (size_t)n_envs * 4) == cudaSuccess
        && cudaMalloc(&b->rlist, (size_t)n_envs * 4) == cudaSuccess
        && cudaMalloc(&b->rcount, 4) == cudaSuccess
        && cudaMalloc(&b->amask, n_envs) == cudaSuccess
        && cudaMallocHost(&b->host_blk, ENV_STRIDE) == cudaSuccess
        && cudaMallocHost(&b->host_i32, (size_t)n_envs * 16) == cudaSuccess;
    if (ok) {
        b->sort_tmp_bytes = 0;
        ok = cub::DeviceRadixSort::SortPairs(NULL, b->sort_tmp_bytes, b->skey, b->skey_out, b->sval, b->perm, n_envs, 0, 36) == cudaSuccess
            && cudaMalloc(&b->sort_tmp, b->sort_tmp_bytes ? b->sort_tmp_bytes : 16) == cudaSuccess;
    }
    if (!ok) { mpb_destroy(b); return NULL; }
    cudaMemcpy(b->anim, anim, sizeof(anim), cudaMemcpyHostToDevice);
    cudaMemset(b->cycles, 0, (size_t)n_envs * 8);
    cudaMemset(b->cost, 0, (size_t)n_envs * 4);

    b->use_sched = 1;
    // Env groups on streams (the heavy 1/64 + two more, 50 frames per launch) pay off while a launch has only a few
    // waves of CTAs, or CTAs of very different length: small batches (2 048 envs +4 %, 8 192 envs +2 %) and batches
    // whose envs are in many schedule classes (staggered episodes: 23.8 ms per step against 27.4).  A single-class
    // batch that gives every SM several 32-env CTAs per launch runs as ONE group with a whole simTick per launch
    // (32 768 envs: 1.76 M env-ticks/s against 1.70 M).  step_groups() / step_frames() choose per step.
    b->groups = 3;
    b->big_groups = 1; b->big_frames = 100;
    *b->ncls_host = 1;
    if (getenv("MPB_FRAMES_PER_LAUNCH")) b->big_frames = b->frames_per_launch;
    if (const char *gg = getenv("MPB_GROUPS")) { int v = atoi(gg); if (v >= 1 && v <= 8) { b->groups = v; b->big_groups = v; } }
    for (int g = 0; g < 8; g++) { cudaStreamCreateWithFlags(&b->gstream[g], cudaStreamNonBlocking); cudaEventCreate(&b->ev_done[g]); }
    cudaEventCreate(&b->ev_start);
    cudaEventCreate(&b->ev_ar0); cudaEventCreate(&b->ev_ar1);
    b->timing = getenv("MPB_TIMING") != NULL;
    b->resched_mid = getenv("MPB_RESCHED_MID") ? atoi(getenv("MPB_RESCHED_MID")) : 1;
    // 1: the tool-ordered action kernel of one-group steps, 2: the action kernels of multi-group steps as well (a few
    // thousand envs in all are better off with a warp each: 2 048 envs on the 120x100 window 331 k against 328 k)
    b->pre_t = getenv("MPB_PRE_T") ? atoi(getenv("MPB_PRE_T")) : PRE_T;
    b->pre_l = getenv("MPB_PRE_L") ? atoi(getenv("MPB_PRE_L")) : PRE_L;
    b->pre_t = (b->pre_t < 32 ? 32 : (b->pre_t > 256 ? 256 : b->pre_t)) & ~31;       // whole warps, the kernel's launch bound
    b->pre_l = b->pre_l < 1 ? 1 : (b->pre_l > 32 ? 32 : b->pre_l);
    b->pre_simt = getenv("MPB_PRE_SIMT") ? atoi(getenv("MPB_PRE_SIMT")) : (n_envs >= 4096 ? 2 : 1);

    for (int i = 0; i < 6; i++) cudaEventCreate(&b->ev_ph[i]);
    if (const char *sc = getenv("MPB_SCHED")) b->use_sched = atoi(sc) != 0;
    k_mp_sched_identity<<<(n_envs + 255) / 256, 256>>>(b->perm, n_envs);
    k_mp_construct<<<(n_envs + MP_EPC - 1) / MP_EPC, MP_THREADS, MP_DYN_SMEM>>>(b->blocks, b->anim, n_envs);
    if (cudaDeviceSynchronize() != cudaSuccess) { mpb_destroy(b); return NULL; }
    return b;
}

void mpb_destroy(mpb_handle h)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return;
    cudaSetDevice(b->device);
    for (int g = 0; g < 8; g++) { if (b->gstream[g]) cudaStreamDestroy(b->gstream[g]); if (b->ev_done[g]) cudaEventDestroy(b->ev_done[g]); }
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    if (b->ev_ar0) cudaEventDestroy(b->ev_ar0);
    if (b->ev_ar1) cudaEventDestroy(b->ev_ar1);
    for (int i = 0; i < 6; i++) if (b->ev_ph[i]) cudaEventDestroy(b->ev_ph[i]);
    cudaFree(b->cycles); cudaFree(b->cost); cudaFree(b->perm); cudaFree(b->cls_sum); cudaFree(b->cls_cnt);
    cudaFree(b->skey); cudaFree(b->skey_out); cudaFree(b->sval); cudaFree(b->sort_tmp); cudaFree(b->ncls_dev); cudaFree(b->perm_pre);
    if (b->ncls_host) cudaFreeHost(b->ncls_host);
    cudaFree(b->prof); cudaFree(b->xt_rnd); cudaFree(b->counters_dev);
    cudaFree(b->rlist); cudaFree(b->rcount); cudaFree(b->amask);
    if (b->pf_stream) { cudaStreamSynchronize(b->pf_stream); cudaStreamDestroy(b->pf_stream); }
    if (b->ev_pf) cudaEventDestroy(b->ev_pf);
    cudaFree(b->next_map); cudaFree(b->next_ep); cudaFree(b->pf_stats);
    cudaFree(b->blocks); cudaFree(b->anim); cudaFree(b->i32_a); cudaFree(b->i32_b); cudaFree(b->mask);
    cudaFree(b->shadow); cudaFree(b->obs); cudaFree(b->reward); cudaFree(b->done); cudaFree(b->act_dev);
    if (b->host_blk) cudaFreeHost(b->host_blk);
    if (b->host_i32) cudaFreeHost(b->host_i32);
    if (b->h_act) cudaFreeHost(b->h_act);
    if (b->h_reward) cudaFreeHost(b->h_reward);
    if (b->h_done) cudaFreeHost(b->h_done);
    if (b->host_shadow) cudaFreeHost(b->host_shadow);
    free(b);
}

const char *mpb_last_error(mpb_handle h) { return h ? ((MpBatch *)h)->err : "null handle"; }
int mpb_num_envs(mpb_handle h) { return h ? ((MpBatch *)h)->n_envs : -1; }
int mpb_env_stride(void) { return ENV_STRIDE; }
int mpb_launches_per_tick(mpb_handle h)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    int per = b->passes > 0 ? (b->passes + b->frames_per_launch - 1) / b->frames_per_launch : 0;
    return 2 * per + (b->passes > 0 ? 0 : 1);
}
int mpb_launches_per_env_step(mpb_handle h)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    int total = b->use_sched ? 2 : 0;                       // the two scheduling kernels of this library (the sort is CUB's)
    if (b->use_sched && step_groups(b) == 1 && b->resched_mid) total += 2;      // ... again between the two simTicks
    if (b->auto_reset) total += 5 + (b->next_map ? 1 : 0);  // list, reset, 2 x frames, post, terrain prefetch
    const int G = step_groups(b);
    if (G == 1 && b->use_sched) total += 1;                 // k_mp_pre_keys (the action kernel runs in tool order)
    for (int g = 0; g < G; g++) {
        const GroupGeom q = group_geom(b, G, g);
        if (q.count <= 0) continue;
        const int F = (G > 1 && g == 0 && q.stride == 1 && q.chunk == 1) ? (b->passes > 0 ? b->passes : 1) : step_frames(b);
        const int per = b->passes > 0 ? (b->passes + F - 1) / F : 1;
        total += 2 + 2 * per;
    }
    return total;
}
int mpb_state_bytes(void) { return EnvLayout::PERSISTENT_END; }
void *mpb_blocks(mpb_handle h) { return h ? ((MpBatch *)h)->blocks : NULL; }

int mpb_seed(mpb_handle h, const int32_t *seeds_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !seeds_host) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->host_i32, seeds_host, (size_t)b->n_envs * 4);
    CK(cudaMemcpyAsync(b->i32_a, b->host_i32, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
    k_mp_set_scalar<<<(b->n_envs + 127) / 128, 128, 0, st>>>(b->blocks, b->n_envs, 0, b->i32_a, 0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

static int set_all(MpBatch *b, int which, int64_t v, cudaStream_t st)
{
    k_mp_set_scalar<<<(b->n_envs + 127) / 128, 128, 0, st>>>(b->blocks, b->n_envs, which, NULL, v);
    CK(cudaGetLastError());
    return 0;
}
int mpb_set_funds(mpb_handle h, int funds, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    return set_all(b, 1, funds, (cudaStream_t)stream);
}
int mpb_set_passes(mpb_handle h, int passes, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    b->passes = passes;
    return set_all(b, 2, passes, (cudaStream_t)stream);
}
int mpb_set_disasters(mpb_handle h, int enabled, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    return set_all(b, 3, enabled ? 1 : 0, (cudaStream_t)stream);
}

static int upload_mask(MpBatch *b, const uint8_t *mask_host, cudaStream_t st)
{
    if (!mask_host) return 0;
    CK(cudaMemcpyAsync(b->mask, mask_host, b->n_envs, cudaMemcpyHostToDevice, st));
    return 0;
}

int mpb_clear_map(mpb_handle h, const uint8_t *mask_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (upload_mask(b, mask_host, st)) return -1;
    k_mp_reset<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->n_envs, NULL, NULL, mask_host ? b->mask : NULL, 0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_generate(mpb_handle h, const int32_t *terrain_seeds_host, const int32_t *sim_seeds_host,
                 const uint8_t *mask_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !terrain_seeds_host || !sim_seeds_host) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (upload_mask(b, mask_host, st)) return -1;
    memcpy(b->host_i32, terrain_seeds_host, (size_t)b->n_envs * 4);
    memcpy(b->host_i32 + b->n_envs, sim_seeds_host, (size_t)b->n_envs * 4);
    CK(cudaMemcpyAsync(b->i32_a, b->host_i32, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b->i32_b, b->host_i32 + b->n_envs, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
    k_mp_reset<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->n_envs, b->i32_a, b->i32_b,
                                                mask_host ? b->mask : NULL, 1);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_tool(mpb_handle h, const int32_t *tool_xy_dev, int32_t *results_dev, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !tool_xy_dev) return -1;
    cudaSetDevice(b->device);
    k_mp_tool<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, (cudaStream_t)stream>>>(b->blocks, b->anim, b->n_envs, tool_xy_dev, results_dev);
    CK(cudaGetLastError());
    return 0;
}

int mpb_tool_host(mpb_handle h, const int32_t *tool_xy_host, int32_t *results_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !tool_xy_host) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->host_i32, tool_xy_host, (size_t)b->n_envs * 12);
    CK(cudaMemcpyAsync(b->i32_a, b->host_i32, (size_t)b->n_envs * 12, cudaMemcpyHostToDevice, st));
    if (mpb_tool(h, b->i32_a, b->i32_b, stream)) return -1;
    CK(cudaMemcpyAsync(b->host_i32, b->i32_b, (size_t)b->n_envs * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (results_host) memcpy(results_host, b->host_i32, (size_t)b->n_envs * 4);
    return 0;
}

int mpb_sim_frames(mpb_handle h, int n, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || n < 0) return -1;
    cudaSetDevice(b->device);
    const int F = b->frames_per_launch;
    for (int f = 0; f < n; f += F)
        launch_frames(b, b->fepc, (cudaStream_t)stream, 0, (n - f < F) ? n - f : F, 0, 0, b->use_sched ? b->perm : NULL,
                      GroupGeom{0, 1, b->n_envs, 1}, NULL);
    CK_LAUNCHES();
    return 0;
}

int mpb_control_sim_tick(mpb_handle h, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    launch_control_sim_tick(b, (cudaStream_t)stream, NULL);
    CK_LAUNCHES();
    return 0;
}

int mpb_sim_tick(mpb_handle h, int animate, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    launch_sim_tick(b, (cudaStream_t)stream, 0, NULL);
    if (animate)
        launch_frames(b, b->fepc, (cudaStream_t)stream, 0, 0, 1, 1, b->use_sched ? b->perm : NULL, GroupGeom{0, 1, b->n_envs, 1}, NULL);
    CK_LAUNCHES();
    return 0;
}

int mpb_call(mpb_handle h, int fn, int a, int bb, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    k_mp_call<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, (cudaStream_t)stream>>>(b->blocks, b->anim, b->n_envs, fn, a, bb);
    CK(cudaGetLastError());
    return 0;
}

int mpb_read_cycles(mpb_handle h, unsigned long long *out_host)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !out_host) return -1;
    cudaSetDevice(b->device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out_host, b->cycles, (size_t)b->n_envs * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemset(b->cycles, 0, (size_t)b->n_envs * 8));
    CK(cudaDeviceSynchronize());       // the legacy stream does not order this against work on non-blocking streams
    return 0;
}

int mpb_group_times(mpb_handle h, float *ms_out)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !ms_out) return -1;
    cudaSetDevice(b->device);
    CK(cudaDeviceSynchronize());
    for (int g = 0; g < b->groups; g++) {
        ms_out[g] = -1.0f;
        if (b->groups > 1) cudaEventElapsedTime(&ms_out[g], b->ev_start, b->ev_done[g]);
    }
    cudaGetLastError();
    return b->groups;
}

// Profiling (MPB_TIMING=1, one-group steps): device time of the phases of the last env step -- schedule rebuild,
// action kernel (with its tool-order sort), the two simTicks, observation / reward, auto-reset -- float [5] ms.
int mpb_step_phase_times(mpb_handle h, float *ms_out)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !ms_out || !b->timing) return -1;
    cudaSetDevice(b->device);
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < 5; i++)
        if (cudaEventElapsedTime(&ms_out[i], b->ev_ph[i], b->ev_ph[i + 1]) != cudaSuccess) { cudaGetLastError(); ms_out[i] = -1.0f; }
    return 0;
}

int mpb_auto_reset_time(mpb_handle h, float *ms_out, int *n_reset_out)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !ms_out) return -1;
    cudaSetDevice(b->device);
    CK(cudaDeviceSynchronize());
    *ms_out = -1.0f;
    if (cudaEventElapsedTime(ms_out, b->ev_ar0, b->ev_ar1) != cudaSuccess) { cudaGetLastError(); *ms_out = -1.0f; }
    if (n_reset_out) CK(cudaMemcpy(n_reset_out, b->rcount, 4, cudaMemcpyDeviceToHost));
    return 0;
}

// Function-hit coverage (-DMP_COVERAGE builds; mp_cov.h): copies the bitmap (one bit per engine / gym function, set
// when any env entered it since the library was loaded) and returns the number of functions; -1 in a product build.
int mpb_coverage(unsigned long long *bits_out, int cap_words)
{
#if defined(MP_COVERAGE)
    if (!bits_out || cap_words < COV_WORDS) return -1;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(bits_out, mp_cov_bits, sizeof(unsigned long long) * COV_WORDS) != cudaSuccess) return -1;
    return COV_COUNT;
#else
    (void)bits_out; (void)cap_words;
    return -1;
#endif
}
const char *mpb_coverage_name(int id)
{
#define MP_COV_NAME(n) #n,
    static const char *const names[] = { MP_COV_LIST(MP_COV_NAME) };
#undef MP_COV_NAME
    return (id >= 0 && id < COV_COUNT) ? names[id] : "";
}

int mpb_read_schedule(mpb_handle h, int *perm_out_host)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !perm_out_host) return -1;
    cudaSetDevice(b->device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(perm_out_host, b->perm, (size_t)b->n_envs * sizeof(int), cudaMemcpyDeviceToHost));
    return b->groups;
}

int mpb_frame_profile(mpb_handle h, int enable, unsigned int *out_host)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    CK(cudaDeviceSynchronize());
    const size_t bytes = (size_t)b->n_envs * 256 * sizeof(unsigned int);
    if (out_host && b->prof) CK(cudaMemcpy(out_host, b->prof, bytes, cudaMemcpyDeviceToHost));
    if (enable && !b->prof) { CK(cudaMalloc(&b->prof, bytes)); }
    if (enable) { CK(cudaMemset(b->prof, 0, bytes)); CK(cudaDeviceSynchronize()); }
    if (!enable && b->prof) { cudaFree(b->prof); b->prof = NULL; }
    return 0;
}

int mpb_sync(mpb_handle h)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    CK(cudaDeviceSynchronize());
    return 0;
}

int mpb_get_field(mpb_handle h, int env, int field, void *out_host, int cap)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    return snapshot_get_field(e, field, out_host, cap);
}

int mpb_set_field(mpb_handle h, int env, int field, const void *in_host, int nbytes)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    if (snapshot_set_field(e, field, in_host, nbytes)) return -1;
    return store_env(b, env);
}

int mpb_get_scalars(mpb_handle h, int env, mp_scalars *out)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !out) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    snapshot_get_scalars(e, out);
    return 0;
}

int mpb_set_scalars(mpb_handle h, int env, const mp_scalars *in)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !in) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    snapshot_set_scalars(e, in);
    return store_env(b, env);
}

int mpb_get_sprites(mpb_handle h, int env, mp_sprite *out, int cap)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !out) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    return snapshot_get_sprites(e, out, cap);
}


// ------------------------------------------------------------------ gym layer (mpb_env_*)
int mpb_env_configure(mpb_handle h, int W, int H, int xs, int ys, int num_tools, const int8_t *tool_zone,
                      const int8_t *tool_engine, int max_step, const double *trg, const double *weight)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || W < 1 || H < 1 || W > 120 || H > 100 || num_tools < 1 || num_tools > GYM_MAX_TOOLS
        || !tool_zone || !tool_engine || !trg || !weight) {
        if (b) snprintf(b->err, sizeof(b->err), "mpb_env_configure: bad arguments");
        return -1;
    }
    cudaSetDevice(b->device);
    cudaFree(b->shadow); cudaFree(b->obs); cudaFree(b->reward); cudaFree(b->done); cudaFree(b->act_dev);
    b->shadow = NULL; b->obs = NULL; b->reward = NULL; b->done = NULL; b->act_dev = NULL;
    GymCfg &c = b->cfg;
    memset(&c, 0, sizeof(c));
    c.W = W; c.H = H; c.xs = xs; c.ys = ys; c.num_tools = num_tools; c.max_step = max_step;
    c.init_funds = 2000000;
    for (int i = 0; i < num_tools; i++) { c.tool_zone[i] = tool_zone[i]; c.tool_engine[i] = tool_engine[i]; }
    for (int i = 0; i < GYM_NUM_METRICS; i++) { c.trg[i] = trg[i]; c.weight[i] = weight[i]; }
    b->shadow_stride = (gym_shadow_bytes(W, H) + 127) & ~127;
    size_t on = (size_t)GYM_OBS_CHANNELS * W * H;
    CK(cudaMalloc(&b->shadow, (size_t)b->n_envs * b->shadow_stride));
    CK(cudaMalloc(&b->obs, (size_t)b->n_envs * on * 4));
    CK(cudaMalloc(&b->reward, (size_t)b->n_envs * 4));
    CK(cudaMalloc(&b->done, b->n_envs));
    CK(cudaMalloc(&b->act_dev, (size_t)b->n_envs * 8));
    if (!b->h_act) {
        CK(cudaMallocHost(&b->h_act, (size_t)b->n_envs * 8));
        CK(cudaMallocHost(&b->h_reward, (size_t)b->n_envs * 4));
        CK(cudaMallocHost(&b->h_done, b->n_envs));
    }
    if (b->host_shadow) cudaFreeHost(b->host_shadow);
    CK(cudaMallocHost(&b->host_shadow, b->shadow_stride));
    CK(cudaMemset(b->shadow, 0, (size_t)b->n_envs * b->shadow_stride));
    CK(cudaMemset(b->obs, 0, (size_t)b->n_envs * on * 4));
    b->obs_out = b->obs;
    CK(cudaMemset(b->reward, 0, (size_t)b->n_envs * 4));
    CK(cudaMemset(b->done, 0, b->n_envs));
    b->configured = 1;
    // MicropolisControl.__init__ builds an empty TileMap (tilemap.py:185 setEmpty)
    k_mp_env_reset_begin<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg,
                                                   NULL, NULL, NULL, 0, 0, 0, NULL, NULL, NULL);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

#define NEED_CFG() do { if (!b || !b->configured) { if (b) snprintf(b->err, sizeof(b->err), "mpb_env_configure not called"); return -1; } cudaSetDevice(b->device); } while (0)

int mpb_env_set_targets(mpb_handle h, const double *trg, const double *weight)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    for (int i = 0; i < GYM_NUM_METRICS; i++) {
        if (trg) b->cfg.trg[i] = trg[i];
        if (weight) b->cfg.weight[i] = weight[i];
    }
    return 0;
}

int mpb_env_set_max_step(mpb_handle h, int max_step)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    b->cfg.max_step = max_step;
    return 0;
}

int mpb_env_set_seed(mpb_handle h, long long seed, int env_offset)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (b->next_ep && (b->run_seed != seed || b->env_offset != env_offset)) {      // prepared maps belong to the old seeds
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(b->next_ep, 0xff, (size_t)b->n_envs * 4));
        CK(cudaDeviceSynchronize());   // memsets run on the legacy stream: callers' non-blocking streams are not ordered after it
    }
    b->run_seed = seed; b->env_offset = env_offset;
    return 0;
}

int mpb_env_update_map(mpb_handle h, const uint8_t *mask_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    cudaStream_t st = (cudaStream_t)stream;
    if (upload_mask(b, mask_host, st)) return -1;
    k_mp_env_update_map<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg,
                                                                   mask_host ? b->mask : NULL);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_env_set_num_step(mpb_handle h, const int32_t *num_step_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (!num_step_host) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->host_i32, num_step_host, (size_t)b->n_envs * 4);
    CK(cudaMemcpyAsync(b->i32_a, b->host_i32, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
    k_mp_env_set_num_step<<<(b->n_envs + 127) / 128, 128, 0, st>>>(b->shadow, b->shadow_stride, b->cfg, b->n_envs, b->i32_a);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_env_set_auto_reset(mpb_handle h, int enable, int terrain)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    b->auto_reset = enable ? 1 : 0; b->auto_terrain = terrain ? 1 : 0;
    if (b->auto_reset && b->auto_terrain && !b->next_map && !getenv("MPB_NO_PREFETCH")) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);         // lo = the lowest priority
        CK(cudaMalloc(&b->next_map, (size_t)b->n_envs * NTILES * sizeof(uint16_t)));
        CK(cudaMalloc(&b->next_ep, (size_t)b->n_envs * 4));
        CK(cudaMalloc(&b->pf_stats, 8));
        CK(cudaMemset(b->next_ep, 0xff, (size_t)b->n_envs * 4));
        CK(cudaMemset(b->pf_stats, 0, 8));
        CK(cudaDeviceSynchronize());   // (a step on a non-blocking stream must not read next_ep before it is cleared)
        CK(cudaStreamCreateWithPriority(&b->pf_stream, cudaStreamNonBlocking, lo));
        CK(cudaEventCreateWithFlags(&b->ev_pf, cudaEventDisableTiming));
    }
    return 0;
}

int mpb_env_prefetch_stats(mpb_handle h, unsigned int *served_out, unsigned int *inline_out)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    unsigned int v[2] = { 0, 0 };
    CK(cudaDeviceSynchronize());
    if (b->pf_stats) CK(cudaMemcpy(v, b->pf_stats, 8, cudaMemcpyDeviceToHost));
    if (served_out) *served_out = v[0];
    if (inline_out) *inline_out = v[1];
    return b->next_map ? 1 : 0;
}

int mpb_env_reset_begin(mpb_handle h, int terrain, const int32_t *terrain_seeds_host,
                        const int32_t *sim_seeds_host, const uint8_t *mask_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    cudaStream_t st = (cudaStream_t)stream;
    const bool host_seeds = terrain && terrain_seeds_host && sim_seeds_host;
    if (terrain && !host_seeds && (terrain_seeds_host || sim_seeds_host)) return -1;     // both or neither
    if (upload_mask(b, mask_host, st)) return -1;
    if (host_seeds) {
        memcpy(b->host_i32, terrain_seeds_host, (size_t)b->n_envs * 4);
        memcpy(b->host_i32 + b->n_envs, sim_seeds_host, (size_t)b->n_envs * 4);
        CK(cudaMemcpyAsync(b->i32_a, b->host_i32, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b->i32_b, b->host_i32 + b->n_envs, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
    }
    k_mp_env_reset_begin<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride,
        b->cfg, host_seeds ? b->i32_a : NULL, host_seeds ? b->i32_b : NULL, mask_host ? b->mask : NULL, terrain,
        b->run_seed, b->env_offset, b->next_map, b->next_ep, b->pf_stats);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_env_reset_finish(mpb_handle h, const uint8_t *mask_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    cudaStream_t st = (cudaStream_t)stream;
    if (upload_mask(b, mask_host, st)) return -1;
    launch_control_sim_tick(b, st, mask_host ? b->mask : NULL);
    k_mp_env_post<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg,
                                                    b->obs_out, b->reward, b->done, 1, mask_host ? b->mask : NULL, NULL, GroupGeom{0, 1, b->n_envs, 1});
    CK_LAUNCHES();
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_env_action(mpb_handle h, const int64_t *actions_host, int static_build, int full_tools,
                   int32_t *ok_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (!actions_host) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->h_act, actions_host, (size_t)b->n_envs * 8);
    CK(cudaMemcpyAsync(b->act_dev, b->h_act, (size_t)b->n_envs * 8, cudaMemcpyHostToDevice, st));
    k_mp_env_action<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg,
                                             b->act_dev, static_build, full_tools, b->i32_b);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(b->host_i32, b->i32_b, (size_t)b->n_envs * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ok_host) memcpy(ok_host, b->host_i32, (size_t)b->n_envs * 4);
    return 0;
}

int mpb_env_set_poet(mpb_handle h, int enable, const double *param_ranges)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (enable && !param_ranges) return -1;
    b->cfg.poet = enable ? 1 : 0;
    for (int i = 0; i < GYM_NUM_METRICS; i++) b->cfg.param_range[i] = enable ? param_ranges[i] : 0.0;
    const size_t on = (size_t)gym_obs_channels(b->cfg) * b->cfg.W * b->cfg.H;
    CK(cudaDeviceSynchronize());
    cudaFree(b->obs);
    b->obs = NULL;
    CK(cudaMalloc(&b->obs, (size_t)b->n_envs * on * 4));
    CK(cudaMemset(b->obs, 0, (size_t)b->n_envs * on * 4));
    CK(cudaDeviceSynchronize());
    b->obs_out = b->obs;
    return gym_obs_channels(b->cfg);
}

int mpb_env_set_obs_buffer(mpb_handle h, float *obs_dev)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    b->obs_out = obs_dev ? obs_dev : b->obs;
    return 0;
}

int mpb_env_track_ages(mpb_handle h, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    cudaStream_t st = (cudaStream_t)stream;
    k_mp_env_track_ages<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

static int env_extinguish(mpb_handle h, int type, int n_dels, const int32_t *rnd_host, const uint8_t *mask_host,
                          int32_t *dels_host, double interval, void *stream);
int mpb_env_extinguish(mpb_handle h, int type, int n_dels, const int32_t *rnd_host, const uint8_t *mask_host,
                       int32_t *dels_host, void *stream)
{
    return env_extinguish(h, type, n_dels, rnd_host, mask_host, dels_host, 0.0, stream);
}
int mpb_env_extinguish_due(mpb_handle h, int type, int n_dels, const int32_t *rnd_host, double interval, int32_t *dels_host,
                           void *stream)
{
    if (interval == 0.0) return -1;
    return env_extinguish(h, type, n_dels, rnd_host, NULL, dels_host, interval, stream);
}
static int env_extinguish(mpb_handle h, int type, int n_dels, const int32_t *rnd_host, const uint8_t *mask_host,
                          int32_t *dels_host, double interval, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (type < 0 || type > 3 || n_dels < 0 || !rnd_host) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t rb = (size_t)b->n_envs * (n_dels + 2) * sizeof(int32_t);
    if (rb > b->xt_rnd_bytes) {
        CK(cudaStreamSynchronize(st));
        cudaFree(b->xt_rnd);
        b->xt_rnd = NULL; b->xt_rnd_bytes = 0;
        CK(cudaMalloc(&b->xt_rnd, rb));
        b->xt_rnd_bytes = rb;
    }
    int32_t *rnd_dev = b->xt_rnd;
    CK(cudaMemcpyAsync(rnd_dev, rnd_host, rb, cudaMemcpyHostToDevice, st));
    if (upload_mask(b, mask_host, st)) return -1;
    k_mp_env_extinguish<<<MP_GRID(b), MP_THREADS, MP_DYN_SMEM, st>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg,
                                                          type, n_dels, rnd_dev, mask_host ? b->mask : NULL, b->i32_b, interval);
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess) le = cudaMemcpyAsync(b->host_i32, b->i32_b, (size_t)b->n_envs * 4, cudaMemcpyDeviceToHost, st);
    if (le == cudaSuccess) le = cudaStreamSynchronize(st);
    if (le != cudaSuccess) return fail(b, "mpb_env_extinguish", le);
    if (dels_host) memcpy(dels_host, b->host_i32, (size_t)b->n_envs * 4);
    return 0;
}

int mpb_env_get_ages(mpb_handle h, int env, int32_t *age_out)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (env < 0 || env >= b->n_envs || !age_out) return -1;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(b->host_shadow, b->shadow + (size_t)env * b->shadow_stride, b->shadow_stride, cudaMemcpyDeviceToHost));
    Gym gy;
    bind_gym(gy, b->host_shadow, &b->cfg);
    memcpy(age_out, gy.age, (size_t)b->cfg.W * b->cfg.H * 4);
    return 0;
}

static int env_step_impl(MpBatch *b, const int64_t *actions_dev, const uint8_t *paint_dev, int static_build,
                         const uint8_t *mask_dev, cudaStream_t st);

int mpb_env_step(mpb_handle h, const int64_t *actions_dev, int static_build, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (!actions_dev) return -1;
    return env_step_impl(b, actions_dev, NULL, static_build, NULL, (cudaStream_t)stream);
}

int mpb_env_step_masked(mpb_handle h, const int64_t *actions_host, int static_build, const uint8_t *mask_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (!actions_host || !mask_host) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->h_act, actions_host, (size_t)b->n_envs * 8);
    CK(cudaMemcpyAsync(b->act_dev, b->h_act, (size_t)b->n_envs * 8, cudaMemcpyHostToDevice, st));
    if (upload_mask(b, mask_host, st)) return -1;
    const int rc = env_step_impl(b, b->act_dev, NULL, static_build, b->mask, st);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_env_step_paint(mpb_handle h, const uint8_t *tools_dev, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (!tools_dev) return -1;
    return env_step_impl(b, NULL, tools_dev, 0, NULL, (cudaStream_t)stream);
}

static int env_step_impl(MpBatch *b, const int64_t *actions_dev, const uint8_t *paint_dev, int static_build,
                         const uint8_t *mask_dev, cudaStream_t st)
{
    const bool big = single_class_big(b);
    const int G = big ? b->big_groups : b->groups;
    b->cur_frames = big ? b->big_frames : b->frames_per_launch;
    const int *perm = b->use_sched ? b->perm : NULL;
    // the order of this step: by the phase each env is in NOW (resets between steps move envs to another phase
    // class) and by what the previous step cost it
    const bool tm = b->timing && G == 1;
    if (tm) cudaEventRecord(b->ev_ph[0], st);
    rebuild_schedule(b, st);
    if (tm) cudaEventRecord(b->ev_ph[1], st);
    if (G > 1) cudaEventRecord(b->ev_start, st);
    for (int g = 0; g < G; g++) {
        cudaStream_t gs = G > 1 ? b->gstream[g] : st;
        const GroupGeom q = group_geom(b, G, g);
        const int grid = MP_GRID_N(q.count);
        if (grid <= 0) continue;
        if (G > 1) cudaStreamWaitEvent(gs, b->ev_start, 0);
        if (paint_dev && b->pre_simt >= 2)
            k_mp_env_pre_t<<<(q.count + b->pre_t / 32 * b->pre_l - 1) / (b->pre_t / 32 * b->pre_l), b->pre_t, b->pre_t / 32 * b->pre_l * SCAL_BYTES, gs>>>(
                b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg, NULL, paint_dev, 0, perm, q, NULL, b->pre_l);
        else if (paint_dev)
            k_mp_env_pre_paint<<<grid, MP_THREADS, MP_DYN_SMEM, gs>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride,
                                                           b->cfg, paint_dev, perm, q);
        else if (G == 1 && b->use_sched && !getenv("MPB_NO_PRE_SORT")) {
            // one group: the action kernel takes the envs in tool order (its own permutation; the frames keep theirs)
            const int n = b->n_envs;
            unsigned int *k32 = reinterpret_cast<unsigned int *>(b->skey), *k32o = reinterpret_cast<unsigned int *>(b->skey_out);
            k_mp_pre_keys<<<(n + 255) / 256, 256, 0, gs>>>(actions_dev, n, b->cfg.W * b->cfg.H, k32, b->sval);
            size_t bytes = b->sort_tmp_bytes;
            const cudaError_t se = cub::DeviceRadixSort::SortPairs(b->sort_tmp, bytes, k32, k32o, b->sval, b->perm_pre, n, 0, 5, gs);
            if (se != cudaSuccess && b->launch_err == cudaSuccess) b->launch_err = se;
            if (b->pre_simt)
                k_mp_env_pre_t<<<(q.count + b->pre_t / 32 * b->pre_l - 1) / (b->pre_t / 32 * b->pre_l), b->pre_t, b->pre_t / 32 * b->pre_l * SCAL_BYTES, gs>>>(
                    b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg, actions_dev, NULL, static_build, b->perm_pre,
                    q, mask_dev, b->pre_l);
            else
                k_mp_env_pre<<<grid, MP_THREADS, MP_DYN_SMEM, gs>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg,
                                                         actions_dev, static_build, b->perm_pre, q, mask_dev);
        } else if (b->pre_simt >= 2)
            k_mp_env_pre_t<<<(q.count + b->pre_t / 32 * b->pre_l - 1) / (b->pre_t / 32 * b->pre_l), b->pre_t, b->pre_t / 32 * b->pre_l * SCAL_BYTES, gs>>>(
                b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg, actions_dev, NULL, static_build, perm, q, mask_dev, b->pre_l);
        else
            k_mp_env_pre<<<grid, MP_THREADS, MP_DYN_SMEM, gs>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg,
                                                     actions_dev, static_build, perm, q, mask_dev);
        if (tm) cudaEventRecord(b->ev_ph[2], gs);
        launch_control_sim_tick(b, gs, mask_dev, G, g, G == 1 && b->use_sched && b->resched_mid);
        if (tm) cudaEventRecord(b->ev_ph[3], gs);
        k_mp_env_post<<<grid, MP_THREADS, MP_DYN_SMEM, gs>>>(b->blocks, b->anim, b->n_envs, b->shadow, b->shadow_stride, b->cfg,
                                                  b->obs_out, b->reward, b->done, 0, mask_dev, perm, q);
        if (tm) cudaEventRecord(b->ev_ph[4], gs);
        if (G > 1) {
            cudaEventRecord(b->ev_done[g], gs);
            cudaStreamWaitEvent(st, b->ev_done[g], 0);
        }
    }
    b->cur_frames = 0;
    if (b->auto_reset) launch_auto_reset(b, mask_dev, st);
    if (tm) cudaEventRecord(b->ev_ph[5], st);
    CK_LAUNCHES();
    return 0;
}

int mpb_env_step_host(mpb_handle h, const int64_t *actions_host, float *reward_host, uint8_t *done_host,
                      void *stream)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (!actions_host || !reward_host || !done_host) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->h_act, actions_host, (size_t)b->n_envs * 8);
    CK(cudaMemcpyAsync(b->act_dev, b->h_act, (size_t)b->n_envs * 8, cudaMemcpyHostToDevice, st));
    if (mpb_env_step(h, b->act_dev, 0, stream)) return -1;
    CK(cudaMemcpyAsync(b->h_reward, b->reward, (size_t)b->n_envs * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(b->h_done, b->done, b->n_envs, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(reward_host, b->h_reward, (size_t)b->n_envs * 4);
    memcpy(done_host, b->h_done, b->n_envs);
    return 0;
}

int mpb_env_get_counters(mpb_handle h, double *out_host)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (!out_host) return -1;
    size_t bytes = (size_t)b->n_envs * MPB_ENV_NUM_COUNTERS * sizeof(double);
    if (!b->counters_dev) CK(cudaMalloc(&b->counters_dev, bytes));
    CK(cudaDeviceSynchronize());
    k_mp_env_counters<<<(b->n_envs + 127) / 128, 128>>>(b->blocks, b->shadow, b->shadow_stride, b->cfg, b->n_envs, b->counters_dev);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out_host, b->counters_dev, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

void *mpb_env_ptr(mpb_handle h, int which)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !b->configured) return NULL;
    switch (which) {
    case MPB_ENV_OBS: return b->obs;
    case MPB_ENV_REWARD: return b->reward;
    case MPB_ENV_DONE: return b->done;
    default: return NULL;
    }
}

int mpb_env_get_shadow(mpb_handle h, int env, uint8_t *zone_out, uint8_t *static_out, int16_t *center_out,
                       double *counters_out)
{
    MpBatch *b = (MpBatch *)h;
    NEED_CFG();
    if (env < 0 || env >= b->n_envs) return -1;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(b->host_shadow, b->shadow + (size_t)env * b->shadow_stride, b->shadow_stride,
                  cudaMemcpyDeviceToHost));
    Gym gy;
    bind_gym(gy, b->host_shadow, &b->cfg);
    int n = b->cfg.W * b->cfg.H;
    if (zone_out) memcpy(zone_out, gy.zone, n);
    if (static_out) memcpy(static_out, gy.stat, n);
    if (center_out) memcpy(center_out, gy.center, (size_t)n * 2);
    if (counters_out) {
        const GymScalars &g = *gy.g;
        double v[MPB_ENV_NUM_COUNTERS] = { (double)g.num_plants, (double)g.num_roads, (double)g.n_struct_tiles,
            (double)g.total_traffic, (double)g.num_step, (double)g.tilemap_error, g.reward, (double)g.done,
            (double)g.curr_funds, g.metrics[0], g.metrics[1], g.metrics[2], g.metrics[3], g.metrics[4],
            g.metrics[5], 0.0, (double)g.episode };
        {   // sprite-table overflow flag of the engine (Scalars::spriteOverflow)
            Env e;
            if (fetch_env(b, env, e)) return -1;
            v[15] = (double)e.s->spriteOverflow;
        }
        memcpy(counters_out, v, sizeof(v));
    }
    return 0;
}

} // extern "C"
