// host_engine.cpp -- TEST INFRASTRUCTURE ONLY (CPU CI harness; never loaded by gymcity_b200).
//
// Compiles the SAME engine headers the CUDA kernels are built from (gymcity_b200/csrc/mp_*.h) with
// g++ for one host thread, behind the `mph_*` mirror of the oracle shim's ABI (oracle/ref_shim.cpp),
// so the CPU test suite can diff the engine logic against the reference engine function by
// function and over long rollouts without a GPU.  The product path is the CUDA library only.
#include <stdlib.h>
#include <string.h>
#include "../../gymcity_b200/csrc/mp_env.h"
#include "../../gymcity_b200/csrc/mp_gym.h"
#include "../../include/mp_snapshot.h"
#include "../../gymcity_b200/csrc/mp_snapshot_io.h"

using namespace mp;

struct HostEnv {
    uint8_t *blk;
    Env e;
    int red[8];
    uint16_t anim[1024];
    GymCfg cfg;
    uint8_t *shadow;
    Gym gy;
};

extern "C" {

void *mph_new(void)
{
    HostEnv *h = new HostEnv();
    h->shadow = nullptr;
    h->blk = (uint8_t *)aligned_alloc(64, (EnvLayout::TOTAL + 63) & ~63);
    bind_env(h->e, h->blk);
    h->e.c.red = h->red;
    for (int i = 0; i < 1024; i++) h->anim[i] = (uint16_t)i;
    for (int i = 0; i < kAnimPairCount; i++) h->anim[kAnimPairs[i][0]] = kAnimPairs[i][1];
    h->e.anim = h->anim;
    env_construct(h->e, h->blk);
    return h;
}
void mph_free(void *p) { HostEnv *h = (HostEnv *)p; free(h->blk); free(h->shadow); delete h; }
void mph_seed(void *p, int seed) { ((HostEnv *)p)->e.s->rng = (uint32_t)seed; }
void mph_clear_map(void *p) { clear_map(((HostEnv *)p)->e); }
void mph_generate(void *p, int seed)
{
    HostEnv *h = (HostEnv *)p;
    generate_some_city(h->e, seed, (int)h->e.s->rng);
}
int mph_tool(void *p, int tool, int x, int y) { return tool_down(((HostEnv *)p)->e, tool, x, y); }
void mph_set_funds(void *p, int f) { ((HostEnv *)p)->e.s->totalFunds = f; }
void mph_set_passes(void *p, int n) { ((HostEnv *)p)->e.s->simPasses = n; }
void mph_sim_tick(void *p) { sim_tick(((HostEnv *)p)->e); }
void mph_tick_engine(void *p)
{
    Env &e = ((HostEnv *)p)->e;
    sim_tick(e);
    if (e.s->doAnimation && !e.s->tilesAnimated) animate_tiles(e);
}
void mph_control_sim_tick(void *p) { control_sim_tick(((HostEnv *)p)->e); }
void mph_sim_frames(void *p, int n) { Env &e = ((HostEnv *)p)->e; for (int i = 0; i < n; i++) sim_loop(e); }
void mph_make_monster(void *p) { make_monster(((HostEnv *)p)->e); }
void mph_call(void *p, int fn, int a, int b) { env_call(((HostEnv *)p)->e, fn, a, b); }
int mph_get_field(void *p, int field, void *out, int cap) { return snapshot_get_field(((HostEnv *)p)->e, field, out, cap); }
int mph_set_field(void *p, int field, const void *in, int n) { return snapshot_set_field(((HostEnv *)p)->e, field, in, n); }
void mph_get_scalars(void *p, mp_scalars *o) { snapshot_get_scalars(((HostEnv *)p)->e, o); }
void mph_set_scalars(void *p, const mp_scalars *o) { snapshot_set_scalars(((HostEnv *)p)->e, o); }
int mph_get_sprites(void *p, mp_sprite *out, int cap) { return snapshot_get_sprites(((HostEnv *)p)->e, out, cap); }


// ---- gym layer (mirrors mpb_env_* for one env)
int mph_env_configure(void *p, int W, int H, int xs, int ys, int num_tools, const int8_t *tool_zone,
                      const int8_t *tool_engine, int max_step, const double *trg, const double *weight)
{
    HostEnv *h = (HostEnv *)p;
    GymCfg &c = h->cfg;
    memset(&c, 0, sizeof(c));
    c.W = W; c.H = H; c.xs = xs; c.ys = ys; c.num_tools = num_tools; c.max_step = max_step; c.init_funds = 2000000;
    for (int i = 0; i < num_tools; i++) { c.tool_zone[i] = tool_zone[i]; c.tool_engine[i] = tool_engine[i]; }
    for (int i = 0; i < GYM_NUM_METRICS; i++) { c.trg[i] = trg[i]; c.weight[i] = weight[i]; }
    free(h->shadow);
    h->shadow = (uint8_t *)calloc(1, gym_shadow_bytes(W, H));
    bind_gym(h->gy, h->shadow, &h->cfg);
    gym_set_empty(h->e, h->gy);
    return 0;
}
void mph_env_reset_begin(void *p, int terrain, int tseed, int sseed)
{
    HostEnv *h = (HostEnv *)p;
    gym_reset_begin(h->e, h->gy, terrain, tseed, sseed);
}
void mph_env_reset_finish(void *p, float *obs, float *reward, uint8_t *done)
{
    HostEnv *h = (HostEnv *)p;
    gym_reset_finish(h->e, h->gy, obs, reward, done);
}
int mph_env_action(void *p, long long a, int static_build, int full)
{
    HostEnv *h = (HostEnv *)p;
    return gym_apply_action(h->e, h->gy, a, static_build != 0, full != 0);
}
void mph_env_step(void *p, long long a, int static_build, float *obs, float *reward, uint8_t *done)
{
    HostEnv *h = (HostEnv *)p;
    gym_step(h->e, h->gy, a, static_build != 0, obs, reward, done);
}
void mph_env_get_shadow(void *p, uint8_t *zone, uint8_t *stat, int16_t *center, double *counters)
{
    HostEnv *h = (HostEnv *)p;
    int n = h->cfg.W * h->cfg.H;
    memcpy(zone, h->gy.zone, n);
    memcpy(stat, h->gy.stat, n);
    memcpy(center, h->gy.center, n * 2);
    const GymScalars &g = *h->gy.g;
    double v[15] = { (double)g.num_plants, (double)g.num_roads, (double)g.n_struct_tiles, (double)g.total_traffic,
        (double)g.num_step, (double)g.tilemap_error, g.reward, (double)g.done, (double)g.curr_funds,
        g.metrics[0], g.metrics[1], g.metrics[2], g.metrics[3], g.metrics[4], g.metrics[5] };
    memcpy(counters, v, sizeof(v));
}

void mph_env_set_poet(void *p, int enable, const double *ranges)
{
    HostEnv *h = (HostEnv *)p;
    h->cfg.poet = enable ? 1 : 0;
    for (int i = 0; i < GYM_NUM_METRICS; i++) h->cfg.param_range[i] = enable ? ranges[i] : 0.0;
}
void mph_env_set_targets(void *p, const double *trg)
{
    HostEnv *h = (HostEnv *)p;
    for (int i = 0; i < GYM_NUM_METRICS; i++) h->cfg.trg[i] = trg[i];
}

// MicropolisPaintEnv.step on the harness
void mph_env_step_paint(void *p, const uint8_t *tools, float *obs, float *reward, uint8_t *done)
{
    HostEnv *h = (HostEnv *)p;
    gym_paint_action(h->e, h->gy, tools);
    h->e.s->totalFunds = h->cfg.init_funds;
    control_sim_tick(h->e);
    gym_step_post(h->e, h->gy, obs, reward, done);
}

// Extinguisher wrapper (wrappers.py) on the harness
void mph_env_track_ages(void *p) { HostEnv *h = (HostEnv *)p; gym_init_age_array(h->e, h->gy); }
int mph_env_extinguish(void *p, int type, int n_dels, const int32_t *rnd)
{
    HostEnv *h = (HostEnv *)p;
    return gym_extinguish(h->e, h->gy, type, n_dels, rnd);
}
void mph_env_get_ages(void *p, int32_t *out)
{
    HostEnv *h = (HostEnv *)p;
    memcpy(out, h->gy.age, (size_t)h->cfg.W * h->cfg.H * 4);
}

} // extern "C"
