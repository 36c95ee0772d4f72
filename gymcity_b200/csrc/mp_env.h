// mp_env.h -- env lifecycle and the env-tick, composed from mp_sim/mp_sprites/mp_tools.
// Shared by the CUDA kernels (mp_batch.cu) and the CPU test harness (tests/host_engine.cpp).
#pragma once
#include "mp_sim.h"
#include "mp_sprites.h"
#include "mp_tools.h"

namespace mp {
#define S (*e.s)

// [ALL] State right after the reference's construction sequence as the gym path runs it:
// Micropolis() -> simInit (main.cpp:135-168) -> initGame (stubs.cpp:240-248) -> cityTax = 10
// (gtkFrontend/main.py:73) -> MicropolisControl.__init__ (corecontrol.py:116-129): setGameLevel(2),
// autoBulldoze, playCity, resume, setFunds(2 000 000), setSpeed(3), setPasses(100).
MPD void env_construct(Env &e, uint8_t *blk)
{
    uint32_t *w = (uint32_t *)blk;
    for (int i = e.c.tid; i < EnvLayout::TOTAL / 4; i += e.c.n) w[i] = 0;
    {   // the scalar block the caller works on (the device kernels stage it in shared memory)
        uint32_t *sw = (uint32_t *)(Scalars *)e.s;
        for (int i = e.c.tid; i < (int)(sizeof(Scalars) / 4); i += e.c.n) sw[i] = 0;
    }
    e.c.sync();
    if (e.c.lead()) {
        S.cityTime = 50;
        S.roadEffect = MAX_ROAD_EFFECT; S.policeEffect = MAX_POLICE_EFFECT; S.fireEffect = MAX_FIRE_EFFECT;
        S.cityTax = 10; S.cityPop = -1; S.cityScore = 500; S.rng = 1; S.totalFunds = 2000000;
        S.autoBulldoze = 1; S.autoBudget = 1; S.gameLevel = 2; S.initSimLoad = 2; S.simSpeed = 3;
        S.simPasses = 100; S.enableDisasters = 1; S.doAnimation = 1;
        S.roadPercent = 1.0f; S.policePercent = 1.0f; S.firePercent = 1.0f; S.externalMarket = 4.0f;
        S.spriteHead = -1;
        MP_NOUNROLL for (int i = 0; i < SPR_COUNT; i++) S.globalSprites[i] = -1;
    }
    e.c.sync();
}

// [ALL] initialize.cpp:76-116 initWillStuff (the time-based reseed is replaced by the explicit
// seed the caller applies afterwards)
MPD void init_will_stuff(Env &e)
{
    if (e.c.lead()) {
        MP_NOUNROLL for (int i = S.spriteHead; i >= 0; i = e.spr()[i].next) e.spr()[i].frame = 0;   // destroyAllSprites
        S.roadEffect = MAX_ROAD_EFFECT; S.policeEffect = MAX_POLICE_EFFECT; S.fireEffect = MAX_FIRE_EFFECT;
        S.cityScore = 500; S.cityPop = -1;
        S.roadFund = 0; S.policeFund = 0; S.fireFund = 0;
        S.disasterEvent = 0; S.taxFlag = 0;
    }
    for (int i = e.c.tid; i < N2; i += e.c.n) {
        e.pop()[i] = 0; e.traffic()[i] = 0; e.pollution()[i] = 0; e.landval()[i] = 0; e.crime()[i] = 0;
    }
    for (int i = e.c.tid; i < N4; i += e.c.n) e.terrain()[i] = 0;
    for (int i = e.c.tid; i < N8; i += e.c.n) {
        e.rog()[i] = 0; e.comRate()[i] = 0; e.policeSt()[i] = 0; e.policeEff()[i] = 0; e.fireSt()[i] = 0;
        e.fireEff()[i] = 0;
    }
    e.c.sync();
}

// [ALL] generate.cpp:165-174 clearMap
MPD void clear_map(Env &e)
{
    for (int i = e.c.tid; i < NTILES; i += e.c.n) e.map()[i] = T_DIRT;
    for (int i = e.c.tid; i < POWER_WORDS; i += e.c.n) { e.act()[i] = 0; e.need()[i] = 0; }
    e.c.sync();
}

// [ALL] generate.cpp:92-110 generateSomeCity(seed), then seedRandom(sim_seed).  `prepared`: the map generateMap(seed)
// produces, made ahead of time (terrain prefetch of the auto-reset, mp_batch.cu) -- copied instead of generated; the
// generator's RNG state does not outlive this function (the simulation seed replaces it below).
MPD void generate_some_city(Env &e, int terrain_seed, int sim_seed, const uint16_t *prepared = nullptr)
{
#if MP_WARP_COOP
    if (prepared) {
        const uint4 *src = reinterpret_cast<const uint4 *>(prepared);
        uint4 *dst = reinterpret_cast<uint4 *>(e.map());
        for (int i = e.c.tid; i < NTILES * 2 / 16; i += e.c.n) dst[i] = src[i];
    } else {
        generate_map_coop(e, terrain_seed);
    }
#else
    (void)prepared;
    if (e.c.lead()) generate_map(e, terrain_seed);
#endif
    e.c.sync();
    rebuild_active(e);              // the terrain generator writes tiles in bulk
    if (e.c.lead()) {
        S.scenario = 0; S.cityTime = 0; S.initSimLoad = 2; S.doInitialEval = 0;
    }
    e.c.sync();
    init_will_stuff(e);
    do_sim_init(e);
    if (e.c.lead()) {
        S.tilesAnimated = 0;               // simUpdate
        S.rng = (uint32_t)sim_seed;
    }
    e.c.sync();
}

// [ALL] C++ simTick (main.cpp:435-443): simPasses x simLoop, then simUpdate (UI flags only)
MPD void sim_tick(Env &e)
{
    if (S.simSpeed) {
        const int n = S.simPasses;
        MP_NOUNROLL for (int i = 0; i < n; i++) sim_loop(e);
    }
}

// [ALL] MicropolisControl.simTick (corecontrol.py:148-152) = tickEngine
// (micropolisgtkengine.py:542-552: simTick, animateTiles, simUpdate) + engine.simTick()
MPD void control_sim_tick(Env &e)
{
    sim_tick(e);
    if (S.doAnimation && !S.tilesAnimated) animate_tiles(e);
    sim_tick(e);
}

// Single-function entry points for injection tests (same ids as oracle/ref_shim.cpp FN_*)
enum Fn : int {
    FN_MAP_SCAN = 0, FN_POWER_SCAN, FN_PTLV_SCAN, FN_CRIME_SCAN, FN_POP_DENSITY_SCAN, FN_FIRE_ANALYSIS,
    FN_DEC_TRAFFIC, FN_DEC_ROG, FN_SET_VALVES, FN_CLEAR_CENSUS, FN_TAKE10, FN_TAKE120, FN_COLLECT_TAX,
    FN_CITY_EVAL, FN_SEND_MESSAGES, FN_DO_DISASTERS, FN_MOVE_OBJECTS, FN_ANIMATE, FN_SIMULATE,
    FN_DO_SIM_INIT, FN_MAKE_FLOOD, FN_MAKE_TORNADO, FN_MAKE_EARTHQUAKE, FN_MAKE_MONSTER, FN_SET_FIRE,
    FN_MAKE_EXPLOSION, FN_GENERATE_SHIP, FN_GENERATE_PLANE, FN_GENERATE_COPTER, FN_GENERATE_TRAIN,
    FN_MAKE_MELTDOWN
};

// [ALL]
MPD void env_call(Env &e, int fn, int a, int b)
{
    const bool lead = e.c.lead();
    switch (fn) {
    case FN_MAP_SCAN: map_scan(e, a, b); break;
    case FN_POWER_SCAN: do_power_scan(e, false); break;
    case FN_PTLV_SCAN: ptlv_scan(e); break;
    case FN_CRIME_SCAN: crime_scan(e); break;
    case FN_POP_DENSITY_SCAN: {
        int cx = S.cityCenterX, cy = S.cityCenterY;
        e.c.sync();
        pop_density_scan(e);
        com_rate_map(e, cx, cy);
        break;
    }
    case FN_FIRE_ANALYSIS: fire_analysis(e); break;
    case FN_DEC_TRAFFIC: dec_traffic_map(e); break;
    case FN_DEC_ROG: dec_rog_map(e); break;
    case FN_SET_VALVES: if (lead) set_valves(e); break;
    case FN_CLEAR_CENSUS:
        if (lead) clear_census_scalars(e);
        for (int i = e.c.tid; i < N8; i += e.c.n) { e.fireSt()[i] = 0; e.policeSt()[i] = 0; }
        break;
    case FN_TAKE10: take10_census(e); break;
    case FN_TAKE120: take120_census(e); break;
    case FN_COLLECT_TAX: if (lead) collect_tax(e); break;
    case FN_CITY_EVAL: if (lead) city_evaluation(e); break;
    case FN_SEND_MESSAGES: if (lead) send_messages(e); break;
    case FN_DO_DISASTERS: {
        int quake = 0;
        if (lead) quake = do_disasters(e, true);
        if (coop_bcast(e, quake)) make_earthquake_coop(e);
        break;
    }
    case FN_MOVE_OBJECTS: if (lead) move_objects(e); break;
    case FN_ANIMATE: animate_tiles(e); break;
    case FN_SIMULATE: simulate(e); break;
    case FN_DO_SIM_INIT: do_sim_init(e); break;
    case FN_MAKE_FLOOD: if (lead) make_flood(e); break;
    case FN_MAKE_TORNADO: if (lead) make_tornado(e); break;
    case FN_MAKE_EARTHQUAKE: make_earthquake_coop(e); break;
    case FN_MAKE_MONSTER: if (lead) make_monster(e); break;
    case FN_SET_FIRE: if (lead) set_fire(e); break;
    case FN_MAKE_EXPLOSION: if (lead) make_explosion(e, a, b); break;
    case FN_GENERATE_SHIP: if (lead) generate_ship(e); break;
    case FN_GENERATE_PLANE: if (lead) generate_plane(e, a, b); break;
    case FN_GENERATE_COPTER: if (lead) generate_copter(e, a, b); break;
    case FN_GENERATE_TRAIN: if (lead) generate_train(e, a, b); break;
    case FN_MAKE_MELTDOWN:
        if (lead)
            MP_NOUNROLL for (int x = 0; x < WORLD_W - 1; x++) {
                bool done = false;
                MP_NOUNROLL for (int y = 0; y < WORLD_H - 1; y++)
                    if ((e.map()[tix(x, y)] & LOMASK) == T_NUCLEAR) { do_meltdown(e, x, y); done = true; break; }
                if (done) break;
            }
        break;
    default: break;
    }
    e.c.sync();
}

#undef S
} // namespace mp
