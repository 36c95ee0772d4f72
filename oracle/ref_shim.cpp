// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" shim over the UNMODIFIED reference Micropolis engine
// (/root/reference/gym_city/envs/micropolis/MicropolisCore/src/MicropolisEngine/src/*.cpp),
// compiled by oracle/build_ref.sh into oracle/_ref/libmicropolis_ref.so.  It replaces the
// SWIG module (swig/micropolisengine.i) that the reference's Python layers bind to and adds
// raw state getters/setters plus single-function entry points for injection tests.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the resulting library.  Nothing under gymcity_b200/ does.
//
// Built with -Dprivate=public (this TU only needs it for member access; the engine TUs get
// it too so the class layout is identical everywhere).
#include "stdafx.h"
#include "micropolis.h"
#include "../include/mp_snapshot.h"

#include <string.h>
#include <time.h>

extern "C" {

// ---------------------------------------------------------------- lifecycle
// Mirrors pyMicropolis/gtkFrontend/main.py:68-87 (train) + gym_city/envs/corecontrol.py:38-135
// (MicropolisControl.__init__): construct, initGame, cityTax=10, setGameLevel(2), playCity
// (setSpeed(2), setPasses(1), resume), resume, setFunds(2e6), setSpeed(3), setPasses(100).
void *mpref_new(void)
{
    Micropolis *m = new Micropolis();
    // The history buffers come from malloc() (allocate.cpp:89-95) and are only zeroed by
    // initSimMemory(); on the empty_start path that never runs, so setValves() would read
    // uninitialised heap.  Pin them to zero (SURVEY 0.5 / DESIGN.md "oracle pinning").
    memset(m->resHist, 0, HISTORY_LENGTH);
    memset(m->comHist, 0, HISTORY_LENGTH);
    memset(m->indHist, 0, HISTORY_LENGTH);
    memset(m->moneyHist, 0, HISTORY_LENGTH);
    memset(m->pollutionHist, 0, HISTORY_LENGTH);
    memset(m->crimeHist, 0, HISTORY_LENGTH);
    memset(m->miscHist, 0, MISC_HISTORY_LENGTH);
    m->initGame();
    m->cityTax = 10;
    m->setGameLevel(LEVEL_HARD);
    m->autoBulldoze = true;
    m->setSpeed(2);
    m->setPasses(1);
    m->resume();
    m->resume();
    m->setFunds(2000000);
    m->setSpeed(3);
    m->setPasses(100);
    m->seedRandom(1);
    return m;
}

void mpref_free(void *h) { delete (Micropolis *)h; }

void mpref_seed(void *h, int seed) { ((Micropolis *)h)->seedRandom(seed); }

void mpref_clear_map(void *h) { ((Micropolis *)h)->clearMap(); }

// engine.generateMap() as the gym path calls it is generateSomeCity(getRandom16()); tests call
// the explicit-seed form and then mpref_seed().
void mpref_generate(void *h, int seed) { ((Micropolis *)h)->generateSomeCity(seed); }

int mpref_tool(void *h, int tool, int x, int y)
{
    return (int)((Micropolis *)h)->toolDown((EditingTool)tool, (short)x, (short)y);
}

void mpref_set_funds(void *h, int funds) { ((Micropolis *)h)->setFunds(funds); }

// fileio.cpp:203-247 loadFileDir: the seven history arrays and the map of a .cty file, nothing else
// (loadFile's scalar decoding reads 8-byte longs from 4-byte fields on LP64 and is not used).
int mpref_load_file_dir(void *h, const char *path) { return ((Micropolis *)h)->loadFileDir(path, NULL) ? 1 : 0; }

void mpref_set_passes(void *h, int passes) { ((Micropolis *)h)->setPasses(passes); }

void mpref_sim_tick(void *h) { ((Micropolis *)h)->simTick(); }

// micropolisgtkengine.py:542-552 tickEngine()
void mpref_tick_engine(void *h)
{
    Micropolis *m = (Micropolis *)h;
    m->simTick();
    if (m->doAnimation && !m->tilesAnimated) {
        m->animateTiles();
    }
    m->simUpdate();
}

// corecontrol.py:148-152 MicropolisControl.simTick()
void mpref_control_sim_tick(void *h)
{
    mpref_tick_engine(h);
    ((Micropolis *)h)->simTick();
}

void mpref_sim_frames(void *h, int n)
{
    Micropolis *m = (Micropolis *)h;
    for (int i = 0; i < n; i++) {
        m->simLoop(true);
    }
}

void mpref_make_monster(void *h) { ((Micropolis *)h)->makeMonster(); }

// ---------------------------------------------------------------- single functions
enum {
    FN_MAP_SCAN = 0, FN_POWER_SCAN, FN_PTLV_SCAN, FN_CRIME_SCAN, FN_POP_DENSITY_SCAN,
    FN_FIRE_ANALYSIS, FN_DEC_TRAFFIC, FN_DEC_ROG, FN_SET_VALVES, FN_CLEAR_CENSUS,
    FN_TAKE10, FN_TAKE120, FN_COLLECT_TAX, FN_CITY_EVAL, FN_SEND_MESSAGES,
    FN_DO_DISASTERS, FN_MOVE_OBJECTS, FN_ANIMATE, FN_SIMULATE, FN_DO_SIM_INIT,
    FN_MAKE_FLOOD, FN_MAKE_TORNADO, FN_MAKE_EARTHQUAKE, FN_MAKE_MONSTER, FN_SET_FIRE,
    FN_MAKE_EXPLOSION, FN_GENERATE_SHIP, FN_GENERATE_PLANE, FN_GENERATE_COPTER,
    FN_GENERATE_TRAIN, FN_MAKE_MELTDOWN,
};

void mpref_call(void *h, int fn, int a, int b)
{
    Micropolis *m = (Micropolis *)h;
    switch (fn) {
    case FN_MAP_SCAN: m->mapScan(a, b); break;
    case FN_POWER_SCAN: m->doPowerScan(); break;
    case FN_PTLV_SCAN: m->pollutionTerrainLandValueScan(); break;
    case FN_CRIME_SCAN: m->crimeScan(); break;
    case FN_POP_DENSITY_SCAN: m->populationDensityScan(); break;
    case FN_FIRE_ANALYSIS: m->fireAnalysis(); break;
    case FN_DEC_TRAFFIC: m->decTrafficMap(); break;
    case FN_DEC_ROG: m->decRateOfGrowthMap(); break;
    case FN_SET_VALVES: m->setValves(); break;
    case FN_CLEAR_CENSUS: m->clearCensus(); break;
    case FN_TAKE10: m->take10Census(); break;
    case FN_TAKE120: m->take120Census(); break;
    case FN_COLLECT_TAX: m->collectTax(); break;
    case FN_CITY_EVAL: m->cityEvaluation(); break;
    case FN_SEND_MESSAGES: m->sendMessages(); break;
    case FN_DO_DISASTERS: m->doDisasters(); break;
    case FN_MOVE_OBJECTS: m->moveObjects(); break;
    case FN_ANIMATE: m->animateTiles(); break;
    case FN_SIMULATE: m->simulate(); break;
    case FN_DO_SIM_INIT: m->doSimInit(); break;
    case FN_MAKE_FLOOD: m->makeFlood(); break;
    case FN_MAKE_TORNADO: m->makeTornado(); break;
    case FN_MAKE_EARTHQUAKE: m->makeEarthquake(); break;
    case FN_MAKE_MONSTER: m->makeMonster(); break;
    case FN_SET_FIRE: m->setFire(); break;
    case FN_MAKE_EXPLOSION: m->makeExplosion(a, b); break;
    case FN_GENERATE_SHIP: m->generateShip(); break;
    case FN_GENERATE_PLANE: m->generatePlane(Position(a, b)); break;
    case FN_GENERATE_COPTER: m->generateCopter(Position(a, b)); break;
    case FN_GENERATE_TRAIN: m->generateTrain(a, b); break;
    case FN_MAKE_MELTDOWN: m->makeMeltdown(); break;
    default: break;
    }
}

// ---------------------------------------------------------------- state access
static void *field_ptr(Micropolis *m, int field, int *nbytes)
{
    switch (field) {
    case MPF_MAP: *nbytes = WORLD_W * WORLD_H * 2; return m->mapBase;
    case MPF_POP_DENSITY: *nbytes = 3000; return m->populationDensityMap.getBase();
    case MPF_TRAFFIC_DENSITY: *nbytes = 3000; return m->trafficDensityMap.getBase();
    case MPF_POLLUTION_DENSITY: *nbytes = 3000; return m->pollutionDensityMap.getBase();
    case MPF_LAND_VALUE: *nbytes = 3000; return m->landValueMap.getBase();
    case MPF_CRIME_RATE: *nbytes = 3000; return m->crimeRateMap.getBase();
    case MPF_TERRAIN_DENSITY: *nbytes = 750; return m->terrainDensityMap.getBase();
    case MPF_POWER_GRID: *nbytes = 12000; return m->powerGridMap.getBase();
    case MPF_RATE_OF_GROWTH: *nbytes = 390; return m->rateOfGrowthMap.getBase();
    case MPF_FIRE_STATION: *nbytes = 390; return m->fireStationMap.getBase();
    case MPF_FIRE_STATION_EFFECT: *nbytes = 390; return m->fireStationEffectMap.getBase();
    case MPF_POLICE_STATION: *nbytes = 390; return m->policeStationMap.getBase();
    case MPF_POLICE_STATION_EFFECT: *nbytes = 390; return m->policeStationEffectMap.getBase();
    case MPF_COM_RATE: *nbytes = 390; return m->comRateMap.getBase();
    case MPF_RES_HIST: *nbytes = 480; return m->resHist;
    case MPF_COM_HIST: *nbytes = 480; return m->comHist;
    case MPF_IND_HIST: *nbytes = 480; return m->indHist;
    case MPF_MONEY_HIST: *nbytes = 480; return m->moneyHist;
    case MPF_POLLUTION_HIST: *nbytes = 480; return m->pollutionHist;
    case MPF_CRIME_HIST: *nbytes = 480; return m->crimeHist;
    case MPF_MISC_HIST: *nbytes = 240; return m->miscHist;
    default: *nbytes = 0; return NULL;
    }
}

int mpref_get_field(void *h, int field, void *out, int cap)
{
    Micropolis *m = (Micropolis *)h;
    if (field == MPF_POWER_STACK) {
        int n = POWER_STACK_SIZE * 2 * 2;
        if (cap < n) return -1;
        int16_t *o = (int16_t *)out;
        for (int i = 0; i < POWER_STACK_SIZE; i++) {
            o[2 * i] = (int16_t)m->powerStackXY[i].posX;
            o[2 * i + 1] = (int16_t)m->powerStackXY[i].posY;
        }
        return n;
    }
    int n = 0;
    void *p = field_ptr(m, field, &n);
    if (!p || cap < n) return -1;
    memcpy(out, p, n);
    return n;
}

int mpref_set_field(void *h, int field, const void *in, int nbytes)
{
    Micropolis *m = (Micropolis *)h;
    if (field == MPF_POWER_STACK) {
        const int16_t *o = (const int16_t *)in;
        for (int i = 0; i < POWER_STACK_SIZE && (i * 4 + 4) <= nbytes; i++) {
            m->powerStackXY[i].posX = o[2 * i];
            m->powerStackXY[i].posY = o[2 * i + 1];
        }
        return 0;
    }
    int n = 0;
    void *p = field_ptr(m, field, &n);
    if (!p || nbytes != n) return -1;
    memcpy(p, in, n);
    return 0;
}

static int count_sprites(Micropolis *m)
{
    int n = 0;
    for (SimSprite *s = m->spriteList; s != NULL; s = s->next) n++;
    return n;
}

void mpref_get_scalars(void *h, mp_scalars *o)
{
    Micropolis *m = (Micropolis *)h;
    memset(o, 0, sizeof(*o));
#define G(n) o->n = (int64_t)m->n;
    G(roadTotal) G(railTotal) G(firePop) G(resPop) G(comPop) G(indPop)
    G(totalPop) G(totalPopLast) G(resZonePop) G(comZonePop) G(indZonePop)
    G(totalZonePop) G(hospitalPop) G(churchPop) G(faith) G(stadiumPop)
    G(policeStationPop) G(fireStationPop) G(coalPowerPop) G(nuclearPowerPop)
    G(seaportPop) G(airportPop) G(crimeAverage) G(pollutionAverage)
    G(landValueAverage) G(cityTime)
    G(roadSpend) G(policeSpend) G(fireSpend) G(roadFund) G(policeFund) G(fireFund)
    G(roadEffect) G(policeEffect) G(fireEffect) G(taxFund) G(cityTax) G(taxFlag)
    G(roadValue) G(policeValue) G(fireValue)
    G(needHospital) G(needChurch) G(floodCount)
    G(cityYes) G(cityPop) G(cityPopDelta) G(cityAssessedValue) G(cityClass)
    G(cityScore) G(cityScoreDelta) G(trafficAverage) G(cityPopLast) G(categoryLast)
    G(powerStackPointer)
    G(cityCenterX) G(cityCenterY) G(pollutionMaxX) G(pollutionMaxY)
    G(crimeMaxX) G(crimeMaxY)
    G(crimeRamp) G(pollutionRamp) G(resCap) G(comCap) G(indCap) G(cashFlow)
    G(disasterEvent) G(disasterWait) G(scoreType) G(scoreWait)
    G(poweredZoneCount) G(unpoweredZoneCount) G(newPower) G(cityTaxAverage)
    G(simCycle) G(phaseCycle) G(speedCycle) G(doInitialEval)
    G(resValve) G(comValve) G(indValve)
    G(spriteCycle) G(totalFunds) G(autoBulldoze) G(autoBudget) G(gameLevel)
    G(initSimLoad) G(scenario) G(simSpeed) G(simPasses) G(enableDisasters)
    G(doAnimation) G(tilesAnimated) G(trafMaxX) G(trafMaxY) G(absDist)
#undef G
    // Only bits 8..23 of the LCG state are ever observed (random.cpp:88-92), and the
    // recurrence is closed modulo 2^32, so the low 32 bits are the whole observable state.
    o->nextRandom = (int64_t)(m->nextRandom & 0xffffffffUL);
    o->numSprites = count_sprites(m);
    for (int i = 0; i < MP_PROBNUM; i++) o->problemVotes[i] = m->problemVotes[i];
    for (int i = 0; i < MP_PROBLEM_COMPLAINTS; i++) o->problemOrder[i] = m->problemOrder[i];
    o->roadPercent = m->roadPercent;
    o->policePercent = m->policePercent;
    o->firePercent = m->firePercent;
    o->externalMarket = m->externalMarket;
}

// Writes back the subset of scalars that injection tests need to control.
void mpref_set_scalars(void *h, const mp_scalars *o)
{
    Micropolis *m = (Micropolis *)h;
#define S(n, T) m->n = (T)o->n;
    S(roadTotal, short) S(railTotal, short) S(firePop, short) S(resPop, short)
    S(comPop, short) S(indPop, short) S(totalPop, short) S(totalPopLast, short)
    S(resZonePop, short) S(comZonePop, short) S(indZonePop, short) S(totalZonePop, short)
    S(hospitalPop, short) S(churchPop, short) S(faith, short) S(stadiumPop, short)
    S(policeStationPop, short) S(fireStationPop, short) S(coalPowerPop, short)
    S(nuclearPowerPop, short) S(seaportPop, short) S(airportPop, short)
    S(crimeAverage, short) S(pollutionAverage, short) S(landValueAverage, short)
    S(cityTime, Quad)
    S(roadSpend, Quad) S(policeSpend, Quad) S(fireSpend, Quad) S(roadFund, Quad)
    S(policeFund, Quad) S(fireFund, Quad) S(roadEffect, Quad) S(policeEffect, Quad)
    S(fireEffect, Quad) S(taxFund, Quad) S(cityTax, short) S(taxFlag, bool)
    S(roadValue, Quad) S(policeValue, Quad) S(fireValue, Quad)
    S(needHospital, short) S(needChurch, short) S(floodCount, short)
    S(cityYes, short) S(cityPop, Quad) S(cityPopDelta, Quad) S(cityAssessedValue, Quad)
    S(cityClass, CityClass) S(cityScore, short) S(cityScoreDelta, short)
    S(trafficAverage, short) S(cityPopLast, Quad) S(categoryLast, short)
    S(powerStackPointer, int) S(nextRandom, UQuad)
    S(cityCenterX, short) S(cityCenterY, short) S(pollutionMaxX, short)
    S(pollutionMaxY, short) S(crimeMaxX, short) S(crimeMaxY, short)
    S(crimeRamp, short) S(pollutionRamp, short) S(resCap, bool) S(comCap, bool)
    S(indCap, bool) S(cashFlow, short) S(disasterEvent, Scenario) S(disasterWait, short)
    S(scoreType, Scenario) S(scoreWait, short) S(poweredZoneCount, short)
    S(unpoweredZoneCount, short) S(newPower, bool) S(cityTaxAverage, short)
    S(simCycle, short) S(phaseCycle, short) S(speedCycle, short) S(doInitialEval, bool)
    S(resValve, short) S(comValve, short) S(indValve, short)
    S(spriteCycle, short) S(totalFunds, Quad) S(autoBulldoze, bool) S(autoBudget, bool)
    S(gameLevel, GameLevel) S(initSimLoad, short) S(scenario, Scenario)
    S(simSpeed, short) S(simPasses, int) S(enableDisasters, bool) S(doAnimation, bool)
    S(tilesAnimated, bool) S(trafMaxX, short) S(trafMaxY, short) S(absDist, int)
#undef S
    for (int i = 0; i < MP_PROBNUM; i++) m->problemVotes[i] = (short)o->problemVotes[i];
    for (int i = 0; i < MP_PROBLEM_COMPLAINTS; i++) m->problemOrder[i] = (short)o->problemOrder[i];
    m->roadPercent = o->roadPercent;
    m->policePercent = o->policePercent;
    m->firePercent = o->firePercent;
    m->externalMarket = o->externalMarket;
}

// Sprites in list order (head first).  Returns the number written.
int mpref_get_sprites(void *h, mp_sprite *out, int cap)
{
    Micropolis *m = (Micropolis *)h;
    int n = 0;
    for (SimSprite *s = m->spriteList; s != NULL && n < cap; s = s->next, n++) {
        mp_sprite *o = &out[n];
        o->type = s->type; o->frame = s->frame; o->x = s->x; o->y = s->y;
        o->width = s->width; o->height = s->height;
        o->xOffset = s->xOffset; o->yOffset = s->yOffset;
        o->xHot = s->xHot; o->yHot = s->yHot;
        o->origX = s->origX; o->origY = s->origY; o->destX = s->destX; o->destY = s->destY;
        o->count = s->count; o->soundCount = s->soundCount; o->dir = s->dir;
        o->newDir = s->newDir; o->step = s->step; o->flag = s->flag;
        o->control = s->control; o->turn = s->turn; o->accel = s->accel; o->speed = s->speed;
    }
    return n;
}

// ---------------------------------------------------------------- CPU baseline replay
// One reference env-tick: tool action(s) are applied by the caller through
// mpref_tool; this runs setFunds + MicropolisControl.simTick().  `mpref_bench_rollout` replays
// `steps` env-ticks of uniformly random single-tile/zone tool actions inside the 16x16 (or WxH)
// window entirely in C++ and returns elapsed seconds, for bench.py's cpu_baseline leg.
static inline uint32_t lcg32(uint32_t *s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

double mpref_bench_rollout(void *h, int steps, int win_w, int win_h, int xs, int ys,
                           unsigned act_seed, long *frames_out)
{
    static const int agent_to_engine[19] = {
        0, 1, 2, 3, 4, 7, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19 };
    Micropolis *m = (Micropolis *)h;
    uint32_t s = act_seed * 2654435761u + 12345u;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < steps; i++) {
        int tool = (int)(lcg32(&s) % 20u);
        int x = (int)(lcg32(&s) % (unsigned)win_w);
        int y = (int)(lcg32(&s) % (unsigned)win_h);
        if (tool < 19) {
            // TileMap.clearTile's triple-tool clear on the target tile, then the build
            // (tilemap.py:410-447, 345-377) -- the engine-side cost of one agent action.
            m->toolDown(TOOL_WATER, (short)(x + xs), (short)(y + ys));
            m->toolDown(TOOL_LAND, (short)(x + xs), (short)(y + ys));
            m->toolDown(TOOL_BULLDOZER, (short)(x + xs), (short)(y + ys));
            m->toolDown((EditingTool)agent_to_engine[tool], (short)(x + xs), (short)(y + ys));
        }
        m->setFunds(2000000);
        mpref_control_sim_tick(h);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (frames_out) *frames_out = (long)steps * 2L * m->simPasses;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

} // extern "C"
