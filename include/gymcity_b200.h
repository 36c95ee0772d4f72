/* gymcity_b200.h -- C-ABI of libgymcity_b200.so (sm_100a CUDA), the drop-in boundary for the
 * environment-tick path of smearle/gym-city.
 *
 * The reference has no C plugin ABI for this path: its operator boundary is SWIG over the whole
 * C++ `Micropolis` class (reference .../MicropolisEngine/swig/micropolisengine.i:66-74) consumed by
 * gym_city/envs/corecontrol.py, and plain torch calls for Game of Life
 * (game_of_life/envs/multi_env.py).  This header is what a maintainer binds instead (ctypes stub in
 * INTEGRATION.md).  Conventions: plain pointers and sizes only; int return 0 = ok, negative = error
 * (text via *_last_error); no exceptions cross the boundary; one handle per GPU; a handle is not
 * re-entrant; all device work is enqueued on the caller's stream (`void *stream` is a
 * cudaStream_t / CUstream, NULL = legacy default stream).
 */
#ifndef GYMCITY_B200_H
#define GYMCITY_B200_H

#include <stdint.h>
#include "mp_snapshot.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ DLPack hand-off
 * Wraps a borrowed device buffer as a DLManagedTensor* (DLPack v0.8 struct layout) whose deleter
 * frees only the descriptor.  dtype_code: 0 int, 1 uint, 2 float.  The caller wraps the pointer in
 * a PyCapsule named "dltensor" and calls torch.from_dlpack (reference hand-off point:
 * storage.py:84-94 RolloutStorage.insert; envs.py:363-375 VecPyTorch.step_wait). */
void *gcb_dlpack_wrap(void *dev_ptr, int dtype_code, int bits, int ndim, const int64_t *shape,
                      int device);
/* The same with a release callback: called once, with release_ctx, when the consumer drops the tensor (its storage is
 * freed).  The host side uses it to keep a handle -- and with it the buffer -- alive while views of it exist
 * (gymcity_b200/dlpack.py: close() / garbage collection of an env is deferred until its last view is gone). */
void *gcb_dlpack_wrap_owned(void *dev_ptr, int dtype_code, int bits, int ndim, const int64_t *shape, int device,
                            void (*release)(void *), void *release_ctx);

/* ------------------------------------------------------------------ Game of Life (golb_*) */
typedef void *golb_handle;

enum golb_buffer {
    GOLB_ROWS = 0,   /* uint32 [n_envs * W]   bit y of rows[env*W + x] = cell [x][y]            */
    GOLB_OBS,        /* multi: float32 [n,2,W,W] (multi_env.py:235-240); classic: uint8 [n,1,W,W] */
    GOLB_REWARD,     /* float32 [n]                                                              */
    GOLB_POP,        /* int32 [n] population after the tick                                      */
    GOLB_DONE        /* uint8 [n]                                                                */
};

/* Replaces GoLMultiEnv.configure (game_of_life/envs/multi_env.py:34-129) when classic = 0 and
 * GameOfLifeEnv.configure (game_of_life/envs/env.py:30-97) batched over n_envs when classic = 1.
 * 3 <= width <= 32.  Returns NULL on bad arguments or if no CUDA device `device` exists. */
golb_handle golb_create(int n_envs, int width, int max_step, int classic, int device);
void golb_destroy(golb_handle h);
const char *golb_last_error(golb_handle h);

/* GoLMultiEnv.set_params (multi_env.py:146-160): target population and the parameter range. */
int golb_set_target(golb_handle h, float trg_pop, float range);

/* World.repopulate_cells (world_pytorch.py:75-82 / world.py:87-92): each cell alive w.p.
 * prob_life in [0,1]; also clears step_count and the classic world's stale references. */
int golb_reset(golb_handle h, uint32_t seed, float prob_life, void *stream);

/* Upload / download explicit boards: uint8 [n_envs][W][W] on the HOST (parity tests, fixtures). */
int golb_set_state(golb_handle h, const uint8_t *cells_host, void *stream);
int golb_get_state(golb_handle h, uint8_t *cells_host);

/* GoLMultiEnv.step (multi_env.py:170-232) / GameOfLifeEnv.step (env.py:125-182):
 * actions_dev = int64 [n_envs] DEVICE pointer, flat index a = x*W + y. */
int golb_step(golb_handle h, const int64_t *actions_dev, void *stream);
/* T steps in ONE launch (multi-env handles): actions_dev = int64 [T, n_envs]; tick t's observation goes to
 * obs_seq_dev[t] (float32 [T, n_envs, 2, W, W], both planes) and its reward to reward_seq_dev[t] (float32
 * [T, n_envs]) when given (either may be NULL); the handle's own obs / reward / pop / done hold the last tick's.
 * Same state as T golb_step calls (multi_env.py:170-232 x T).  For scripted / random-action rollouts: at 4 096
 * boards a tick is ~1 us of GPU work and a launch per tick is launch-bound. */
int golb_step_n(golb_handle h, const int64_t *actions_dev, int T, float *obs_seq_dev, float *reward_seq_dev, void *stream);

/* Same, end to end with HOST buffers: copies actions H2D, steps, copies rewards D2H, syncs. */
int golb_step_host(golb_handle h, const int64_t *actions_host, float *reward_host, void *stream);

/* Borrowed device pointers (owned by the handle, valid until golb_destroy). */
void *golb_ptr(golb_handle h, int which);
int golb_step_count(golb_handle h);

/* ------------------------------------------------------------------ Micropolis engine batch (mpb_*)
 * One handle = n_envs independent Micropolis engines resident on one GPU, each in the state the
 * reference engine has after MicropolisControl.__init__ (gym_city/envs/corecontrol.py:38-135):
 * gameLevel 2, cityTax 10, simSpeed 3, simPasses 100, funds 2 000 000, disasters / autoBulldoze /
 * autoBudget on.  World is 120x100 (map_type.h:73,78); coordinates are world coordinates. */
typedef void *mpb_handle;

mpb_handle mpb_create(int n_envs, int device);
void mpb_destroy(mpb_handle h);
const char *mpb_last_error(mpb_handle h);
int mpb_num_envs(mpb_handle h);
int mpb_launches_per_tick(mpb_handle h);
int mpb_launches_per_env_step(mpb_handle h);   /* kernel launches one mpb_env_step enqueues */   /* kernel launches one MicropolisControl.simTick takes */
int mpb_env_stride(void);            /* bytes between consecutive env blocks in HBM            */
int mpb_state_bytes(void);           /* bytes of persistent state per env (without scratch)     */
void *mpb_blocks(mpb_handle h);      /* borrowed device pointer to the env blocks               */
int mpb_sync(mpb_handle h);
/* Profiling: SM cycles each env's warp spent simulating since the last call (uint64 [n_envs] HOST);
 * max / mean shows the load imbalance a launch has to wait out. */
int mpb_read_cycles(mpb_handle h, unsigned long long *out_host);
/* Profiling: milliseconds from the start of the last mpb_env_step to the end of each env group's
 * work (float [8] HOST); returns the number of groups. */
int mpb_group_times(mpb_handle h, float *ms_out);
/* Profiling: per-env per-frame SM cycles of the next env-steps (uint32 [n_envs][256] HOST, frames 0..199 of
 * an env-step used): enable != 0 starts / clears the recording, out_host (may be NULL) receives what was
 * recorded so far, enable == 0 stops and frees it.  One batch per process at a time. */
int mpb_frame_profile(mpb_handle h, int enable, unsigned int *out_host);
/* Test infrastructure: function-hit coverage of a -DMP_COVERAGE build (csrc/mp_cov.h).  mpb_coverage copies the
 * bitmap (bit i = function i was entered by some env since load; uint64 words) and returns the number of functions,
 * or -1 in a product build; mpb_coverage_name(i) names function i. */
int mpb_coverage(unsigned long long *bits_out, int cap_words);
const char *mpb_coverage_name(int id);
/* Profiling (handles created with MPB_TIMING set): device time (ms) of the five phases of the last one-group env step
 * -- schedule rebuild, action kernel, the two simTicks, observation / reward, auto-reset -- float [5]. */
int mpb_step_phase_times(mpb_handle h, float *ms_out);
/* Profiling: device time (ms) of the auto-reset part of the last env step and how many envs it reset. */
int mpb_auto_reset_time(mpb_handle h, float *ms_out, int *n_reset_out);
/* Profiling: the heavy-first env order the NEXT env-step's launches use (int32 [n_envs] HOST); returns the
 * number of env groups. */
int mpb_read_schedule(mpb_handle h, int *perm_out_host);

/* Micropolis::seedRandom per env (random.cpp:171-174); seeds_host = int32 [n_envs] HOST. */
int mpb_seed(mpb_handle h, const int32_t *seeds_host, void *stream);
/* engine.setFunds / setPasses / setEnableDisasters for every env (corecontrol.py:126-129,258). */
int mpb_set_funds(mpb_handle h, int funds, void *stream);
int mpb_set_passes(mpb_handle h, int passes, void *stream);
int mpb_set_disasters(mpb_handle h, int enabled, void *stream);

/* engine.clearMap (generate.cpp:165-174) on the envs whose mask byte is non-zero (NULL = all). */
int mpb_clear_map(mpb_handle h, const uint8_t *mask_host, void *stream);
/* engine.generateSomeCity(terrain_seed) (generate.cpp:92-110) followed by seedRandom(sim_seed):
 * the reference reseeds from the wall clock inside initWillStuff, so a reproducible reset has to
 * name both seeds.  HOST int32 [n_envs] each; mask as above. */
int mpb_generate(mpb_handle h, const int32_t *terrain_seeds_host, const int32_t *sim_seeds_host,
                 const uint8_t *mask_host, void *stream);

/* engine.toolDown(tool, x, y) (tool.cpp:1487-1516) once per env.  tool_xy = int32 [n_envs][3]
 * {EditingTool, x, y}; tool < 0 = no call.  results = int32 [n_envs] ToolResult (-2 no money,
 * -1 need bulldoze, 0 failed, 1 ok; micropolis.h:387-392), may be NULL. */
int mpb_tool(mpb_handle h, const int32_t *tool_xy_dev, int32_t *results_dev, void *stream);
int mpb_tool_host(mpb_handle h, const int32_t *tool_xy_host, int32_t *results_host, void *stream);

/* n x simLoop(true) (main.cpp:403-425). */
int mpb_sim_frames(mpb_handle h, int n, void *stream);
/* C++ simTick (main.cpp:435-443); animate != 0 adds tickEngine's animateTiles
 * (micropolisgtkengine.py:542-552). */
int mpb_sim_tick(mpb_handle h, int animate, void *stream);
/* MicropolisControl.simTick (corecontrol.py:148-152): tickEngine() + engine.simTick(). */
int mpb_control_sim_tick(mpb_handle h, void *stream);
/* One engine function on every env (ids = enum Fn in csrc/mp_env.h, same as the oracle shim's
 * FN_*): injection tests run a single scan / census / disaster function and diff its outputs. */
int mpb_call(mpb_handle h, int fn, int a, int b, void *stream);

/* State access for one env (HOST buffers; field ids = enum mp_field in mp_snapshot.h).  These
 * stand in for the reference's getTile / getPowerGrid / get*MapBuffer getters (stubs.cpp:353-733). */
int mpb_get_field(mpb_handle h, int env, int field, void *out_host, int cap);
int mpb_set_field(mpb_handle h, int env, int field, const void *in_host, int nbytes);
int mpb_get_scalars(mpb_handle h, int env, mp_scalars *out);
int mpb_set_scalars(mpb_handle h, int env, const mp_scalars *in);
int mpb_get_sprites(mpb_handle h, int env, mp_sprite *out, int cap);

/* ------------------------------------------------------------------ gym layer on the batch (mpb_env_*)
 * MicropolisEnv / MicropolisControl / TileMap for every env of the handle, on the device:
 * gym_city/envs/env.py:448-540 (step/postact), :264-292 (reset), :301-343 (observation),
 * :345-387 (reward); corecontrol.py:215-239 (density maps), :296-338 (tools, window offset);
 * tilemap.py:345-541 (addZoneBot, clear-before-build, shadow map).
 *
 * mpb_env_configure: W x H build window at world offset (xs, ys) (reference: 16, 8,
 * corecontrol.py:58-59); num_tools agent tools with tool_zone[i] = TileMap zone id the tool is
 * recorded as (-1 'Nil', -2 'Clear') and tool_engine[i] = engine EditingTool; episode length
 * max_step; reward targets / weights in the order res_pop, com_pop, ind_pop, traffic, num_plants,
 * mayor_rating (env.py:31-56).  Leaves every env with an empty shadow map. */
enum mpb_env_buffer {
    MPB_ENV_OBS = 0,     /* float32 [n_envs, 32 (38 with poet), W, H], channel order of env.py:315-332 */
    MPB_ENV_REWARD,      /* float32 [n_envs]                                                     */
    MPB_ENV_DONE         /* uint8 [n_envs]                                                       */
};
#define MPB_ENV_NUM_COUNTERS 17
int mpb_env_configure(mpb_handle h, int W, int H, int xs, int ys, int num_tools, const int8_t *tool_zone,
                      const int8_t *tool_engine, int max_step, const double *trg, const double *weight);
int mpb_env_set_targets(mpb_handle h, const double *trg, const double *weight);   /* env.py:419-424 */
int mpb_env_set_max_step(mpb_handle h, int max_step);
/* reset, first half (env.py:264-274): TileMap.setEmpty + engine.clearMap + updateMap and, when
 * terrain != 0, engine.generateSomeCity(terrain_seed) + seedRandom(sim_seed) + updateMap.  Set-up
 * builds (powerPuzzle / randomStaticStart) go through mpb_env_action before the second half. */
int mpb_env_reset_begin(mpb_handle h, int terrain, const int32_t *terrain_seeds_host,
                        const int32_t *sim_seeds_host, const uint8_t *mask_host, void *stream);
/* Seeds derived on the device.  The reference seeds its engine from the wall clock (random.cpp:155-164); here every
 * episode of every env has a terrain seed and a simulation seed = hash(seed, env_offset + i, episode_i, which),
 * episode_i = the number of derived-seed resets env i has taken (counter 16).  mpb_env_reset_begin with
 * terrain != 0 and both seed arrays NULL uses them; so does the auto-reset below.  env_offset = global index of
 * env 0 (sharding: results do not depend on the number of GPUs). */
int mpb_env_set_seed(mpb_handle h, long long seed, int env_offset);
/* The vec-env worker's auto-reset (subproc_vec_env.py:16-21: `if done: ob = env.reset()`), on the device: with
 * enable != 0 every mpb_env_step / _step_host / _step_paint / _step_masked ends by resetting the envs whose done
 * flag it set (terrain != 0: clearMap + generateSomeCity with the env's next derived seeds; 0: clearMap only),
 * running the reset's simTick for them and writing their reset observation.  Reward and done stay those of the
 * step that ended the episode.  No host round trip: the finished envs are a device-side list. */
int mpb_env_set_auto_reset(mpb_handle h, int enable, int terrain);
/* With auto-reset on the terrain path, the map of every env's NEXT episode is generated ahead of time by a
 * low-priority kernel in the gaps of the following steps (the seeds are known as soon as the current episode starts),
 * and a reset copies it instead of generating (ENG/generate.cpp:119-160 is 2.5 ms of one warp's time).  Returns 1 if
 * the prefetch is active; served_out / inline_out: resets served from a prepared map / generated in line so far. */
int mpb_env_prefetch_stats(mpb_handle h, unsigned int *served_out, unsigned int *inline_out);
/* MicropolisControl.updateMap (corecontrol.py:193-203) for the envs of the mask (NULL = all): the TileMap shadow of the
 * build window is re-read from the engine map, e.g. after a city file was loaded into the engine (mpb_set_field MPF_MAP,
 * gymcity_b200.cty.load_env). */
int mpb_env_update_map(mpb_handle h, const uint8_t *mask_host, void *stream);
/* MicropolisEnv.num_step of every env (int32 [n_envs] HOST): episodes that end on different steps (in the reference
 * every worker counts its own steps, env.py:448-460; randomStaticStart gives each its own offset, env.py:228-241). */
int mpb_env_set_num_step(mpb_handle h, const int32_t *num_step_host, void *stream);
/* reset, second half (env.py:279-292): simTick, metrics, setFunds, reward, observation. */
int mpb_env_reset_finish(mpb_handle h, const uint8_t *mask_host, void *stream);
/* MicropolisControl.takeAction without a tick: actions_host = int64 [n_envs] flat action index
 * i = z*W*H + x*H + y (env.py:209-219), < 0 = none; full_tools != 0 decodes z against the full
 * 20-tool agent list (corecontrol.py:85-106) whatever the env's own list is (powerPuzzle builds
 * Residential / NuclearPowerPlant in a Wire-only env); ok_host = int32 [n_envs] tool success. */
int mpb_env_action(mpb_handle h, const int64_t *actions_host, int static_build, int full_tools,
                   int32_t *ok_host, void *stream);

/* mpb_env_step for a subset of the batch: envs with mask_host[i] == 0 (uint8 [n_envs] HOST) are left
 * untouched -- no action, no tick, no observation / reward / done update, num_step unchanged.  Actions are
 * HOST int64 [n_envs].  Used by MicropolisEnv.randomStaticStart (env.py:228-241), where every env takes its own
 * number of full static-build steps during reset.  Synchronises the stream before returning. */
int mpb_env_step_masked(mpb_handle h, const int64_t *actions_host, int static_build, const uint8_t *mask_host, void *stream);

/* MicropolisPaintEnv.step (gym_city/envs/paintenv.py:43-52, paintcontrol.py:45-70): one tool per window
 * cell.  tools_dev: uint8 [n_envs][W*H] DEVICE, index into the env's tool list at cell x*H + y.  Cells are
 * visited row by row (the last column is never reached, as in the reference); a cell an earlier build of
 * the same step covers is skipped and is not cleared by later builds (TileMap.acted, tilemap.py:402-404).
 * Then setFunds, MicropolisControl.simTick, observation / reward / done exactly as mpb_env_step. */
int mpb_env_step_paint(mpb_handle h, const uint8_t *tools_dev, void *stream);

/* MicropolisEnv(poet=True) observation (env.py:301-313): the three population planes are divided by their
 * parameter ranges and six planes are added, target k divided by param_ranges[k] (double [6]), before the
 * static-build plane: 38 planes instead of 32.  Reallocates MPB_ENV_OBS (fetch the pointer again) and
 * returns the number of planes; enable == 0 restores the 32-plane layout. */
int mpb_env_set_poet(mpb_handle h, int enable, const double *param_ranges);

/* Where mpb_env_step / mpb_env_step_host / mpb_env_reset_finish write the observations from now on:
 * float32 [n_envs, 32, W, H] DEVICE memory of the caller -- e.g. RolloutStorage.obs[step + 1]
 * (storage.py:84-94), which saves the learner's copy of the batch of observations -- or NULL for the
 * library's own buffer (MPB_ENV_OBS).  The caller keeps the buffer alive until the step has finished. */
int mpb_env_set_obs_buffer(mpb_handle h, float *obs_dev);

/* Extinguisher wrapper (gym_city/wrappers.py:10-160).
 * mpb_env_track_ages: TileMap.init_age_array (tilemap.py:201-206) for every env: from now on the shadow
 *   map records the num_step of every build (age_order) and -1 for natural cells.
 * mpb_env_extinguish: one extinction event per env (mask_host: uint8 [n_envs] or NULL = all).
 *   type 0 'age' (elderCleanse: the n_dels oldest cells are cleared, then setFunds), 1 'spatial'
 *   (localWipe around a random centre), 2 'random' (ranDemolish), 3 'Monster' (engine.makeMonster()).
 *   rnd_host: int32 [n_envs][n_dels + 2], non-negative draws of the caller's RNG (the reference uses
 *   numpy's global stream): 'random' picks aged cell number rnd[i] % count for deletion i, 'spatial'
 *   uses (rnd[0] % W, rnd[1] % H) as the centre.  dels_host (may be NULL): deletions made, int32 [n_envs].
 * mpb_env_get_ages: age_order of one env, int32 [W*H] HOST. */
int mpb_env_track_ages(mpb_handle h, void *stream);
int mpb_env_extinguish(mpb_handle h, int type, int n_dels, const int32_t *rnd_host, const uint8_t *mask_host,
                       int32_t *dels_host, void *stream);
/* Extinguisher.step's own test (wrappers.py:35-40), per env and on the device: env i is hit iff its num_step > 0 and
 * num_step % interval == 0 (interval = 1 / extinction_prob, a double; -1 for prob 0 hits every step, as in the reference).
 * Envs that were just reset (num_step 0) are skipped: in the reference the wrapper runs before the worker's auto-reset
 * replaces the city. */
int mpb_env_extinguish_due(mpb_handle h, int type, int n_dels, const int32_t *rnd_host, double interval, int32_t *dels_host,
                           void *stream);
int mpb_env_get_ages(mpb_handle h, int env, int32_t *age_out);
/* MicropolisEnv.step for every env: one fused launch (action, setFunds, 2 x simPasses simFrames +
 * animateTiles, observation, reward, done).  actions_dev = int64 [n_envs] DEVICE pointer. */
int mpb_env_step(mpb_handle h, const int64_t *actions_dev, int static_build, void *stream);
/* Same end to end with HOST buffers: H2D actions, step, D2H reward + done, sync. */
int mpb_env_step_host(mpb_handle h, const int64_t *actions_host, float *reward_host, uint8_t *done_host,
                      void *stream);
void *mpb_env_ptr(mpb_handle h, int which);
/* Counters of every env in one call: double [n_envs][MPB_ENV_NUM_COUNTERS] HOST, fields as in
 * mpb_env_get_shadow. */
int mpb_env_get_counters(mpb_handle h, double *out_host);
/* Shadow map of one env (HOST buffers, any may be NULL): zone id / static flag per cell [W*H]
 * (index x*H + y), recorded centre cx*128+cy or -1, and counters {num_plants, num_roads,
 * n_struct_tiles, total_traffic, num_step, tilemap_error, reward, done, funds, metrics[6],
 * sprite_overflow}.  tilemap_error and sprite_overflow are sticky divergence flags: the first is raised
 * where the reference's TileMap would have died with a KeyError (tilemap.py:488-491), the second when an
 * env needed more than MAX_SPRITES sprites at once (the reference's list is unbounded); both must be 0 for
 * a rollout to count as bit-exact. */
int mpb_env_get_shadow(mpb_handle h, int env, uint8_t *zone_out, uint8_t *static_out, int16_t *center_out,
                       double *counters_out);

#ifdef __cplusplus
}
#endif
#endif /* GYMCITY_B200_H */
