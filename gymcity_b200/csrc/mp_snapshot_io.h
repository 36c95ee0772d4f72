// mp_snapshot_io.h -- host-side conversion between one env's state block and the flat snapshot
// types of include/mp_snapshot.h (parity tests diff these against the oracle's).
#pragma once
#include <string.h>
#include "mp_state.h"
#include "../../include/mp_snapshot.h"

namespace mp {

inline void *snapshot_field_ptr(Env &e, int field, int *nbytes)
{
    switch (field) {
    case MPF_MAP: *nbytes = NTILES * 2; return e.map();
    case MPF_POP_DENSITY: *nbytes = N2; return e.pop();
    case MPF_TRAFFIC_DENSITY: *nbytes = N2; return e.traffic();
    case MPF_POLLUTION_DENSITY: *nbytes = N2; return e.pollution();
    case MPF_LAND_VALUE: *nbytes = N2; return e.landval();
    case MPF_CRIME_RATE: *nbytes = N2; return e.crime();
    case MPF_TERRAIN_DENSITY: *nbytes = N4; return e.terrain();
    case MPF_RATE_OF_GROWTH: *nbytes = N8 * 2; return e.rog();
    case MPF_FIRE_STATION: *nbytes = N8 * 2; return e.fireSt();
    case MPF_FIRE_STATION_EFFECT: *nbytes = N8 * 2; return e.fireEff();
    case MPF_POLICE_STATION: *nbytes = N8 * 2; return e.policeSt();
    case MPF_POLICE_STATION_EFFECT: *nbytes = N8 * 2; return e.policeEff();
    case MPF_COM_RATE: *nbytes = N8 * 2; return e.comRate();
    case MPF_RES_HIST: *nbytes = 480; return e.resHist();
    case MPF_COM_HIST: *nbytes = 480; return e.comHist();
    case MPF_IND_HIST: *nbytes = 480; return e.indHist();
    case MPF_MONEY_HIST: *nbytes = 480; return e.moneyHist();
    case MPF_POLLUTION_HIST: *nbytes = 480; return e.pollutionHist();
    case MPF_CRIME_HIST: *nbytes = 480; return e.crimeHist();
    case MPF_MISC_HIST: *nbytes = 240; return e.miscHist();
    default: *nbytes = 0; return nullptr;
    }
}

inline int snapshot_get_field(Env &e, int field, void *out, int cap)
{
    if (field == MPF_POWER_GRID) {
        if (cap < NTILES) return -1;
        uint8_t *o = (uint8_t *)out;
        for (int i = 0; i < NTILES; i++) o[i] = (e.power()[i >> 5] >> (i & 31)) & 1u;
        return NTILES;
    }
    e.s->powerQuiet = 0; e.s->powerPure = 0;
    if (field == MPF_POWER_STACK) {
        int n = POWER_STACK_SIZE * 4;
        if (cap < n) return -1;
        int16_t *o = (int16_t *)out;
        for (int i = 0; i < POWER_STACK_SIZE; i++) {
            o[2 * i] = (int16_t)(e.pstack()[i] / WORLD_H);
            o[2 * i + 1] = (int16_t)(e.pstack()[i] % WORLD_H);
        }
        return n;
    }
    int n = 0;
    void *p = snapshot_field_ptr(e, field, &n);
    if (!p || cap < n) return -1;
    memcpy(out, p, n);
    return n;
}

inline void snapshot_rebuild_bitmaps(Env &e)
{
    e.s->powerQuiet = 0; e.s->powerPure = 0;    // state written from outside: the next power scan runs
    for (int w = 0; w < POWER_WORDS; w++) {
        uint32_t m = 0, nd = 0;
        for (int b = 0; b < 32; b++) {
            const int idx = w * 32 + b, v = e.map()[idx];
            if (!tile_is_active(v)) continue;
            m |= 1u << b;
            const int t = v & LOMASK;
            const bool inert = t >= T_POWERBASE && !(v & ZONEBIT) && !(t >= T_RAILBASE && t < T_RESBASE)
                && !(t >= T_SOMETINYEXP && t <= T_LASTTINYEXP);
            bool need = true;
            if (inert) {
                if (!(v & CONDBIT) || !e.s->newPower) need = false;
                else {
                    const bool want = t == T_NUCLEAR || t == T_POWERPLANT || ((e.power()[idx >> 5] >> (idx & 31)) & 1u);
                    need = ((v & PWRBIT) != 0) != want;
                }
            }
            if (need) nd |= 1u << b;
        }
        e.act()[w] = m;
        e.need()[w] = nd;
    }
}

inline int snapshot_set_field(Env &e, int field, const void *in, int nbytes)
{
    if (field == MPF_POWER_GRID) {
        if (nbytes != NTILES) return -1;
        const uint8_t *s = (const uint8_t *)in;
        for (int i = 0; i < POWER_WORDS; i++) e.power()[i] = 0;
        for (int i = 0; i < NTILES; i++) if (s[i]) e.power()[i >> 5] |= 1u << (i & 31);
        snapshot_rebuild_bitmaps(e);
        return 0;
    }
    if (field == MPF_POWER_STACK) {
        const int16_t *s = (const int16_t *)in;
        for (int i = 0; i < POWER_STACK_SIZE && (i * 4 + 4) <= nbytes; i++)
            e.pstack()[i] = (uint16_t)(s[2 * i] * WORLD_H + s[2 * i + 1]);
        return 0;
    }
    int n = 0;
    void *p = snapshot_field_ptr(e, field, &n);
    if (!p || nbytes != n) return -1;
    memcpy(p, in, n);
    if (field == MPF_MAP) snapshot_rebuild_bitmaps(e);      // keep the derived bitmaps in step
    return 0;
}

#define MP_SNAPSHOT_DIRECT(X) \
    X(roadTotal) X(railTotal) X(firePop) X(resPop) X(comPop) X(indPop) \
    X(totalPop) X(totalPopLast) X(resZonePop) X(comZonePop) X(indZonePop) \
    X(totalZonePop) X(hospitalPop) X(churchPop) X(faith) X(stadiumPop) \
    X(policeStationPop) X(fireStationPop) X(coalPowerPop) X(nuclearPowerPop) \
    X(seaportPop) X(airportPop) X(crimeAverage) X(pollutionAverage) \
    X(landValueAverage) X(cityTime) \
    X(roadSpend) X(policeSpend) X(fireSpend) X(roadFund) X(policeFund) X(fireFund) \
    X(roadEffect) X(policeEffect) X(fireEffect) X(taxFund) X(cityTax) X(taxFlag) \
    X(roadValue) X(policeValue) X(fireValue) \
    X(needHospital) X(needChurch) X(floodCount) \
    X(cityYes) X(cityPop) X(cityPopDelta) X(cityAssessedValue) X(cityClass) \
    X(cityScore) X(cityScoreDelta) X(trafficAverage) X(cityPopLast) X(categoryLast) \
    X(powerStackPointer) \
    X(cityCenterX) X(cityCenterY) X(pollutionMaxX) X(pollutionMaxY) \
    X(crimeMaxX) X(crimeMaxY) \
    X(crimeRamp) X(pollutionRamp) X(resCap) X(comCap) X(indCap) X(cashFlow) \
    X(disasterEvent) X(disasterWait) X(scoreType) X(scoreWait) \
    X(poweredZoneCount) X(unpoweredZoneCount) X(newPower) X(cityTaxAverage) \
    X(simCycle) X(phaseCycle) X(speedCycle) X(doInitialEval) \
    X(resValve) X(comValve) X(indValve) \
    X(spriteCycle) X(totalFunds) X(autoBulldoze) X(autoBudget) X(gameLevel) \
    X(initSimLoad) X(scenario) X(simSpeed) X(simPasses) X(enableDisasters) \
    X(doAnimation) X(tilesAnimated) X(trafMaxX) X(trafMaxY) X(absDist)

inline void snapshot_get_scalars(Env &e, mp_scalars *o)
{
    const Scalars &s = *e.s;
    memset(o, 0, sizeof(*o));
#define X(n) o->n = (int64_t)s.n;
    MP_SNAPSHOT_DIRECT(X)
#undef X
    o->nextRandom = (int64_t)s.rng;
    int cnt = 0;
    for (int i = s.spriteHead; i >= 0; i = e.spr()[i].next) cnt++;
    o->numSprites = cnt;
    for (int i = 0; i < MP_PROBNUM; i++) o->problemVotes[i] = s.problemVotes[i];
    for (int i = 0; i < MP_PROBLEM_COMPLAINTS; i++) o->problemOrder[i] = s.problemOrder[i];
    o->roadPercent = s.roadPercent; o->policePercent = s.policePercent;
    o->firePercent = s.firePercent; o->externalMarket = s.externalMarket;
}

template <typename T> inline void snapshot_assign(T &dst, int64_t v) { dst = (T)v; }

inline void snapshot_set_scalars(Env &e, const mp_scalars *o)
{
    Scalars &s = *e.s;
#define X(n) snapshot_assign(s.n, o->n);
    MP_SNAPSHOT_DIRECT(X)
#undef X
    s.rng = (uint32_t)o->nextRandom;
    for (int i = 0; i < MP_PROBNUM; i++) s.problemVotes[i] = (int16_t)o->problemVotes[i];
    for (int i = 0; i < MP_PROBLEM_COMPLAINTS; i++) s.problemOrder[i] = (int16_t)o->problemOrder[i];
    s.roadPercent = o->roadPercent; s.policePercent = o->policePercent;
    s.firePercent = o->firePercent; s.externalMarket = o->externalMarket;
    snapshot_rebuild_bitmaps(e);            // newPower may have changed
}

inline int snapshot_get_sprites(Env &e, mp_sprite *out, int cap)
{
    int n = 0;
    for (int i = e.s->spriteHead; i >= 0 && n < cap; i = e.spr()[i].next, n++) {
        const Sprite &s = e.spr()[i];
        mp_sprite *o = &out[n];
        o->type = s.type; o->frame = s.frame; o->x = s.x; o->y = s.y; o->width = s.width;
        o->height = s.height; o->xOffset = s.xOffset; o->yOffset = s.yOffset; o->xHot = s.xHot;
        o->yHot = s.yHot; o->origX = s.origX; o->origY = s.origY; o->destX = s.destX;
        o->destY = s.destY; o->count = s.count; o->soundCount = s.soundCount; o->dir = s.dir;
        o->newDir = s.newDir; o->step = s.step; o->flag = s.flag; o->control = s.control;
        o->turn = s.turn; o->accel = s.accel; o->speed = s.speed;
    }
    return n;
}

} // namespace mp
