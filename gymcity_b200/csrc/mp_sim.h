// mp_sim.h -- the Micropolis simulation step: 16-phase cycle, map scan, zones, traffic, power
// flood fill, overlay scans, census / valves / tax / evaluation.
//
// Functions marked [ALL] are entered by every thread of the env's CTA (they contain barriers and
// strided parallel loops); functions marked [LEAD] run on the lead thread only -- that is the
// RNG-ordered serial lane (every draw of the per-env simRandom stream happens there, in the
// reference's call order).  Each function cites the reference code whose results it reproduces
// (paths relative to .../MicropolisEngine/src/).
#pragma once
#include "mp_state.h"

namespace mp {

#define S (*e.s)

MPN void city_evaluation(Env &e);                       // mp_eval part below
MPR void move_objects(Env &e);                          // mp_sprites.h
MPN int do_disasters(Env &e, bool defer_quake);         // mp_sprites.h
MPN void make_earthquake_coop(Env &e);                  // mp_sprites.h
MPN void make_explosion_at(Env &e, int x, int y);       // mp_sprites.h
MPN void make_explosion(Env &e, int x, int y);          // mp_sprites.h
MPN void generate_train_make(Env &e, int x, int y);     // mp_sprites.h
MPN void generate_ship(Env &e);                         // mp_sprites.h
MPN void generate_plane(Env &e, int x, int y);          // mp_sprites.h
MPN void generate_copter(Env &e, int x, int y);         // mp_sprites.h
MPD int get_sprite(Env &e, int type);                   // mp_sprites.h
MPD int boat_distance(Env &e, int x, int y);            // mp_sprites.h

// ------------------------------------------------------------------ phase 0
// [LEAD] simulate.cpp:546-670 setValves.  fp32 throughout like the reference, including the
// `resRatio = min(indRatio, ...)` slip at :644.
// num / den for num >= +0 and den > 0.  A zero numerator (empty histories: the usual case in a young city) sends the
// IEEE division to its slow-path subroutine (two calls per setValves in the ncu call counts); the quotient is +0.
MPD float fdiv_pos(float num, float den) { return num == 0.0f ? 0.0f : num / den; }

MPN void set_valves(Env &e)
{
    MP_COV(set_valves);
    MP_TAB int16_t taxTable[21] = { 200, 150, 120, 100, 80, 50, 30, 0, -10, -40, -100,
        -150, -200, -250, -300, -350, -400, -450, -500, -550, -600 };
    MP_TAB float extMarketParamTable[3] = { 1.2f, 1.1f, 0.98f };
    const float birthRate = 0.02f, laborBaseMax = 1.3f, internalMarketDenom = 3.7f;
    const float projectedIndPopMin = 5.0f, resRatioDefault = 1.3f;
    const float resRatioMax = 2, comRatioMax = 2, indRatioMax = 2, taxTableScale = 600;
    int16_t *mh = e.miscHist();
    mh[1] = f2s(S.externalMarket);
    mh[2] = S.resPop; mh[3] = S.comPop; mh[4] = S.indPop;
    mh[5] = S.resValve; mh[6] = S.comValve; mh[7] = S.indValve;
    mh[10] = S.crimeRamp; mh[11] = S.pollutionRamp; mh[12] = S.landValueAverage;
    mh[13] = S.crimeAverage; mh[14] = S.pollutionAverage; mh[15] = (int16_t)S.gameLevel;
    mh[16] = (int16_t)S.cityClass; mh[17] = S.cityScore;

    float normalizedResPop = (float)S.resPop / 8.0f;
    S.totalPopLast = S.totalPop;
    S.totalPop = f2s(normalizedResPop + S.comPop + S.indPop);
    float employment;
    if (S.resPop > 0) employment = fdiv_pos((float)(e.comHist()[1] + e.indHist()[1]), normalizedResPop);
    else employment = 1;
    float migration = normalizedResPop * (employment - 1);
    float births = normalizedResPop * birthRate;
    float projectedResPop = normalizedResPop + migration + births;
    float temp = (float)(e.comHist()[1] + e.indHist()[1]);
    float laborBase;
    if (temp > 0.0f) laborBase = fdiv_pos((float)e.resHist()[1], temp);
    else laborBase = 1;
    laborBase = tclamp(laborBase, 0.0f, laborBaseMax);
    float internalMarket = (float)(normalizedResPop + S.comPop + S.indPop) / internalMarketDenom;
    float projectedComPop = internalMarket * laborBase;
    float projectedIndPop = S.indPop * laborBase * extMarketParamTable[S.gameLevel];
    projectedIndPop = tmax(projectedIndPop, projectedIndPopMin);
    float resRatio, comRatio, indRatio;
    if (normalizedResPop > 0) resRatio = projectedResPop / normalizedResPop;
    else resRatio = resRatioDefault;
    if (S.comPop > 0) comRatio = fdiv_pos(projectedComPop, (float)S.comPop);
    else comRatio = projectedComPop;
    if (S.indPop > 0) indRatio = projectedIndPop / (float)S.indPop;
    else indRatio = projectedIndPop;
    resRatio = tmin(resRatio, resRatioMax);
    comRatio = tmin(comRatio, comRatioMax);
    resRatio = tmin(indRatio, indRatioMax);        // sic
    int16_t z = tmin((int16_t)(S.cityTax + S.gameLevel), (int16_t)20);
    resRatio = (resRatio - 1) * taxTableScale + taxTable[z];
    comRatio = (comRatio - 1) * taxTableScale + taxTable[z];
    indRatio = (indRatio - 1) * taxTableScale + taxTable[z];
    S.resValve = (int16_t)tclamp(S.resValve + f2s(resRatio), -RES_VALVE_RANGE, RES_VALVE_RANGE);
    S.comValve = (int16_t)tclamp(S.comValve + f2s(comRatio), -COM_VALVE_RANGE, COM_VALVE_RANGE);
    S.indValve = (int16_t)tclamp(S.indValve + f2s(indRatio), -IND_VALVE_RANGE, IND_VALVE_RANGE);
    if (S.resCap && S.resValve > 0) S.resValve = 0;
    if (S.comCap && S.comValve > 0) S.comValve = 0;
    if (S.indCap && S.indValve > 0) S.indValve = 0;
}

// [LEAD] simulate.cpp:674-703 clearCensus, scalar part (the two station maps are cleared by ALL)
MPD void clear_census_scalars(Env &e)
{
    MP_COV(clear_census_scalars);
    S.poweredZoneCount = 0; S.unpoweredZoneCount = 0; S.firePop = 0; S.roadTotal = 0;
    S.railTotal = 0; S.resPop = 0; S.comPop = 0; S.indPop = 0; S.resZonePop = 0;
    S.comZonePop = 0; S.indZonePop = 0; S.hospitalPop = 0; S.churchPop = 0;
    S.policeStationPop = 0; S.fireStationPop = 0; S.stadiumPop = 0; S.coalPowerPop = 0;
    S.nuclearPowerPop = 0; S.seaportPop = 0; S.airportPop = 0; S.powerStackPointer = 0;
}

// ------------------------------------------------------------------ power stack / zone power
MPD void push_power_stack(Env &e, int x, int y)          // power.cpp:163-169
{
    MP_COV(push_power_stack);
    if (S.powerStackPointer < POWER_STACK_SIZE - 2) {
        S.powerStackPointer++;
        e.pstack()[S.powerStackPointer] = (uint16_t)tix(x, y);
    }
}

// [LEAD] zone.cpp:419-437 setZonePower
MPP bool set_zone_power(Env &e, int x, int y)
{
    MP_COV(set_zone_power);
    int v = e.map()[tix(x, y)], t = v & LOMASK;
    if (t == T_NUCLEAR || t == T_POWERPLANT || pwr_get(e, x, y)) {
        e.map()[tix(x, y)] = (uint16_t)(v | PWRBIT);          // status bit only: activity unchanged
        return true;
    }
    e.map()[tix(x, y)] = (uint16_t)(v & ~PWRBIT);
    return false;
}

// ------------------------------------------------------------------ traffic (traffic.cpp)
MPD bool road_test(int mv)                                // traffic.cpp:489-503
{
    MP_COV(road_test);
    int t = mv & LOMASK;
    if (t < T_ROADBASE || t > T_LASTRAIL) return false;
    if (t >= T_POWERBASE && t < T_LASTPOWER) return false;
    return true;
}

// [LEAD] traffic.cpp:227-249 findPerimeterRoad
MPD bool find_perimeter_road(Env &e, int &x, int &y)
{
    MP_COV(find_perimeter_road);
    MP_TAB int8_t px[12] = { -1, 0, 1, 2, 2, 2, 1, 0, -1, -2, -2, -2 };
    MP_TAB int8_t py[12] = { -2, -2, -2, -1, 0, 1, 2, 2, 2, 1, 0, -1 };
    MP_NOUNROLL for (int z = 0; z < 12; z++) {
        int tx = x + px[z], ty = y + py[z];
        if (inb(tx, ty) && road_test(e.map()[tix(tx, ty)])) {
            x = tx; y = ty;
            return true;
        }
    }
    return false;
}

MPD int tile_toward(const Env &e, int x, int y, int dir, int dflt)   // traffic.cpp:383-427
{
    MP_COV(tile_toward);
    switch (dir) {
    case D_N: return y > 0 ? (e.map()[tix(x, y - 1)] & LOMASK) : dflt;
    case D_E: return x < WORLD_W - 1 ? (e.map()[tix(x + 1, y)] & LOMASK) : dflt;
    case D_S: return y < WORLD_H - 1 ? (e.map()[tix(x, y + 1)] & LOMASK) : dflt;
    case D_W: return x > 0 ? (e.map()[tix(x - 1, y)] & LOMASK) : dflt;
    default: return dflt;
    }
}

// [LEAD] traffic.cpp:333-376 tryGo
MPD int try_go(Env &e, int x, int y, int dirLast)
{
    MP_COV(try_go);
    int dirs[4], dir = D_N, count = 0;
    MP_NOUNROLL for (int i = 0; i < 4; i++) {
        if (dir != dirLast && road_test(tile_toward(e, x, y, dir, T_DIRT))) {
            dirs[i] = dir;
            count++;
        } else {
            dirs[i] = D_INVALID;
        }
        dir = rot45(dir, 2);
    }
    if (count == 0) return D_INVALID;
    if (count == 1) {
        MP_NOUNROLL for (int i = 0; i < 4; i++) if (dirs[i] != D_INVALID) return dirs[i];
    }
    int i = rnd16(e) & 3;
    while (dirs[i] == D_INVALID) i = (i + 1) & 3;
    return dirs[i];
}

// [LEAD] traffic.cpp:436-480 driveDone
MPD bool drive_done(const Env &e, int x, int y, int dest)
{
    MP_COV(drive_done);
    MP_TAB int lo[3] = { T_COMBASE, T_LHTHR, T_LHTHR };
    MP_TAB int hi[3] = { T_NUCLEAR, T_PORT, T_COMBASE };
    int l = lo[dest], h = hi[dest], z;
    if (y > 0) { z = e.map()[tix(x, y - 1)] & LOMASK; if (z >= l && z <= h) return true; }
    if (x < WORLD_W - 1) { z = e.map()[tix(x + 1, y)] & LOMASK; if (z >= l && z <= h) return true; }
    if (y < WORLD_H - 1) { z = e.map()[tix(x, y + 1)] & LOMASK; if (z >= l && z <= h) return true; }
    if (x > 0) { z = e.map()[tix(x - 1, y)] & LOMASK; if (z >= l && z <= h) return true; }
    return false;
}

// [LEAD] traffic.cpp:287-323 tryDrive
MPD bool try_drive(Env &e, int x, int y, int dest)
{
    MP_COV(try_drive);
    int dirLast = D_INVALID;
    MP_NOUNROLL for (int16_t dist = 0; dist < MAX_TRAFFIC_DISTANCE; dist++) {
        int dir = try_go(e, x, y, dirLast);
        if (dir != D_INVALID) {
            pos_move(x, y, dir);
            dirLast = rot45(dir, 4);
            if (dist & 1) {
                S.curMapStackPointer++;
                S.curMapStack[S.curMapStackPointer] = (int16_t)tix(x, y);
            }
            if (drive_done(e, x, y, dest)) return true;
        } else {
            if (S.curMapStackPointer > 0) {
                S.curMapStackPointer--;
                dist += 3;
            } else {
                return false;
            }
        }
    }
    return false;
}

// [LEAD] traffic.cpp:164-201 addToTrafficDensityMap
MPD void add_to_traffic_density(Env &e)
{
    MP_COV(add_to_traffic_density);
    while (S.curMapStackPointer > 0) {
        int p = S.curMapStack[S.curMapStackPointer];
        S.curMapStackPointer--;
        int x = p / WORLD_H, y = p - x * WORLD_H;
        int t = e.map()[p] & LOMASK;
        if (t >= T_ROADBASE && t < T_POWERBASE) {
            int traffic = b2get(e.traffic(), x, y) + 50;
            traffic = tmin(traffic, 240);
            b2set(e.traffic(), x, y, traffic);
            if (traffic >= 240 && rnd(e, 5) == 0) {
                S.trafMaxX = (int16_t)x; S.trafMaxY = (int16_t)y;
                int sp = get_sprite(e, SPR_HELICOPTER);
                if (sp >= 0 && e.spr()[sp].control == -1) {
                    e.spr()[sp].destX = (int16_t)(S.trafMaxX * 16);
                    e.spr()[sp].destY = (int16_t)(S.trafMaxY * 16);
                }
            }
        }
    }
}

// [LEAD] traffic.cpp:126-156 makeTraffic: 1 = trip found, 0 = no trip, -1 = no road at all
MPN int make_traffic(Env &e, int x, int y, int dest)
{
    MP_COV(make_traffic);
    S.curMapStackPointer = 0;
    if (find_perimeter_road(e, x, y)) {
        if (try_drive(e, x, y, dest)) {
            add_to_traffic_density(e);
            return 1;
        }
        return 0;
    }
    return -1;
}

// ------------------------------------------------------------------ zones (zone.cpp)
MPD void inc_rate_of_growth(Env &e, int x, int y, int amount)      // zone.cpp:300-306
{
    MP_COV(inc_rate_of_growth);
    int v = s8get(e.rog(), x, y);
    v = tclamp(v + amount * 4, -200, 200);
    s8set(e.rog(), x, y, v);
}

MPN int16_t land_pollution_value(const Env &e, int x, int y)       // zone.cpp:271-293
{
    MP_COV(land_pollution_value);
    int16_t lv = (int16_t)b2get(e.landval(), x, y);
    lv = (int16_t)(lv - b2get(e.pollution(), x, y));
    if (lv < 30) return 0;
    if (lv < 80) return 1;
    if (lv < 150) return 2;
    return 3;
}

// [LEAD] zone.cpp:315-352 zonePlop
MPN bool zone_plop(Env &e, int x, int y, int base)
{
    MP_COV(zone_plop);
    // z = 0..8 in the reference's order: x fastest (xx = x + z % 3 - 1, yy = y + z / 3 - 1)
    MP_NOUNROLL for (int yy = y - 1; yy <= y + 1; yy++)
        MP_NOUNROLL for (int xx = x - 1; xx <= x + 1; xx++)
            if (inb(xx, yy)) {
                int t = e.map()[tix(xx, yy)] & LOMASK;
                if (t >= T_FLOOD && t < T_ROADBASE) return false;
            }
    MP_NOUNROLL for (int yy = y - 1; yy <= y + 1; yy++)
        MP_NOUNROLL for (int xx = x - 1; xx <= x + 1; xx++) {
            if (inb(xx, yy)) mset(e, tix(xx, yy), (uint16_t)(base + BNCNBIT));
            base++;
        }
    set_zone_power(e, x, y);
    mset(e, tix(x, y), e.map()[tix(x, y)] | (ZONEBIT + BULLBIT));
    return true;
}

MPD int16_t do_free_pop(const Env &e, int x, int y)                // zone.cpp:360-377
{
    MP_COV(do_free_pop);
    int16_t count = 0;
    MP_NOUNROLL for (int xx = x - 1; xx <= x + 1; xx++)
        MP_NOUNROLL for (int yy = y - 1; yy <= y + 1; yy++)
            if (inb(xx, yy)) {
                int t = e.map()[tix(xx, yy)] & LOMASK;
                if (t >= T_LHTHR && t <= T_HHTHR) count++;
            }
    return count;
}

MPD int16_t res_zone_pop(int t) { return (int16_t)((((t - T_RZB) / 9) % 4) * 8 + 16); }   // zone.cpp:727
MPD int16_t com_zone_pop(int t) { return t == T_COMCLR ? (int16_t)0 : (int16_t)(((t - T_CZB) / 9) % 5 + 1); }
MPD int16_t ind_zone_pop(int t) { return t == T_INDCLR ? (int16_t)0 : (int16_t)((((t - T_IZB) / 9) % 4) + 1); }

MPD int16_t eval_lot(const Env &e, int x, int y)                   // zone.cpp:493-517
{
    MP_COV(eval_lot);
    MP_TAB int8_t dx[4] = { 0, 1, 0, -1 }, dy[4] = { -1, 0, 1, 0 };
    int16_t z = (int16_t)(e.map()[tix(x, y)] & LOMASK);
    if (z && (z < T_RESBASE || z > T_RESBASE + 8)) return -1;
    int16_t score = 1;
    MP_NOUNROLL for (z = 0; z < 4; z++) {
        int xx = x + dx[z], yy = y + dy[z];
        if (inb(xx, yy) && e.map()[tix(xx, yy)] != T_DIRT && (e.map()[tix(xx, yy)] & LOMASK) <= T_LASTROAD)
            score++;
    }
    return score;
}

// [LEAD] zone.cpp:444-484 buildHouse
MPN void build_house(Env &e, int x, int y, int value)
{
    MP_COV(build_house);
    MP_TAB int8_t zx[9] = { 0, -1, 0, 1, -1, 1, -1, 0, 1 }, zy[9] = { 0, -1, -1, -1, 0, 0, 1, 1, 1 };
    int16_t best = 0, hscore = 0;
    MP_NOUNROLL for (int16_t z = 1; z < 9; z++) {
        int xx = x + zx[z], yy = y + zy[z];
        if (inb(xx, yy)) {
            int16_t score = eval_lot(e, xx, yy);
            if (score != 0) {
                if (score > hscore) { hscore = score; best = z; }
                if (score == hscore && !(rnd16(e) & 7)) best = z;
            }
        }
    }
    if (best) {
        int xx = x + zx[best], yy = y + zy[best];
        if (inb(xx, yy)) mset(e, tix(xx, yy), (uint16_t)(T_HOUSE + BLBNCNBIT + rnd(e, 2) + value * 3));
    }
}

MPD void res_plop(Env &e, int x, int y, int den, int value)        // zone.cpp:741-747
{
    MP_COV(res_plop);
    int16_t base = (int16_t)(((value * 4 + den) * 9) + T_RZB - 4);
    zone_plop(e, x, y, base);
}
MPD void com_plop(Env &e, int x, int y, int den, int value)        // zone.cpp:900-906
{
    MP_COV(com_plop);
    int16_t base = (int16_t)(((value * 5) + den) * 9 + T_CZB - 4);
    zone_plop(e, x, y, base);
}
MPD void ind_plop(Env &e, int x, int y, int den, int value)        // zone.cpp:1005-1009
{
    MP_COV(ind_plop);
    int16_t base = (int16_t)(((value * 4) + den) * 9 + T_IND1);
    zone_plop(e, x, y, base);
}

// [LEAD] zone.cpp:596-630 doResIn
MPN void do_res_in(Env &e, int x, int y, int pop, int value)
{
    MP_COV(do_res_in);
    if (b2get(e.pollution(), x, y) > 128) return;
    int t = e.map()[tix(x, y)] & LOMASK;
    if (t == T_FREEZ) {
        if (pop < 8) {
            build_house(e, x, y, value);
            inc_rate_of_growth(e, x, y, 1);
            return;
        }
        if (b2get(e.pop(), x, y) > 64) {
            res_plop(e, x, y, 0, value);
            inc_rate_of_growth(e, x, y, 8);
        }
        return;
    }
    if (pop < 40) {
        res_plop(e, x, y, (pop / 8) - 1, value);
        inc_rate_of_growth(e, x, y, 8);
    }
}

// [LEAD] zone.cpp:639-704 doResOut
MPN void do_res_out(Env &e, int x, int y, int pop, int value)
{
    MP_COV(do_res_out);
    MP_TAB int8_t brdr[9] = { 0, 3, 6, 1, 4, 7, 2, 5, 8 };
    if (!pop) return;
    if (pop > 16) {
        res_plop(e, x, y, (pop - 24) / 8, value);
        inc_rate_of_growth(e, x, y, -8);
        return;
    }
    if (pop == 16) {
        inc_rate_of_growth(e, x, y, -8);
        mset(e, tix(x, y), (uint16_t)(T_FREEZ | BLBNCNBIT | ZONEBIT));
        MP_NOUNROLL for (int xx = x - 1; xx <= x + 1; xx++)
            MP_NOUNROLL for (int yy = y - 1; yy <= y + 1; yy++)
                if (inb(xx, yy) && (e.map()[tix(xx, yy)] & LOMASK) != T_FREEZ)
                    mset(e, tix(xx, yy), (uint16_t)(T_LHTHR + value + rnd(e, 2) + BLBNCNBIT));
    }
    if (pop < 16) {
        inc_rate_of_growth(e, x, y, -1);
        int z = 0;
        MP_NOUNROLL for (int xx = x - 1; xx <= x + 1; xx++)
            MP_NOUNROLL for (int yy = y - 1; yy <= y + 1; yy++) {
                if (inb(xx, yy)) {
                    int loc = e.map()[tix(xx, yy)] & LOMASK;
                    if (loc >= T_LHTHR && loc <= T_HHTHR) {
                        mset(e, tix(xx, yy), (uint16_t)(brdr[z] + BLBNCNBIT + T_FREEZ - 4));
                        return;
                    }
                }
                z++;
            }
    }
}

// [LEAD] zone.cpp:240-264 makeHospital
MPN void make_hospital(Env &e, int x, int y)
{
    MP_COV(make_hospital);
    if (S.needHospital > 0) {
        zone_plop(e, x, y, T_HOSPITAL - 4);
        S.needHospital = 0;
        return;
    }
    if (S.needChurch > 0) {
        int churchType = rnd(e, 7);
        int tile = churchType == 0 ? (int)T_CHURCH0 : T_CHURCH1 + ((churchType - 1) * 9);
        zone_plop(e, x, y, tile - 4);
        S.needChurch = 0;
    }
}

MPD int16_t eval_res(const Env &e, int x, int y, int traf)         // zone.cpp:757-777
{
    MP_COV(eval_res);
    if (traf < 0) return -3000;
    int16_t v = (int16_t)b2get(e.landval(), x, y);
    v = (int16_t)(v - b2get(e.pollution(), x, y));
    if (v < 0) v = 0;
    else v = (int16_t)tmin(v * 32, 6000);
    return (int16_t)(v - 3000);
}

// [LEAD] zone.cpp:526-587 doResidential
MPN void do_residential(Env &e, int x, int y, bool zonePower)
{
    MP_COV(do_residential);
    S.resZonePop++;
    int t = e.map()[tix(x, y)] & LOMASK;
    int16_t tpop = (t == T_FREEZ) ? do_free_pop(e, x, y) : res_zone_pop(t);
    S.resPop = (int16_t)(S.resPop + tpop);
    int16_t trf = (tpop > rnd(e, 35)) ? (int16_t)make_traffic(e, x, y, 0) : (int16_t)1;
    if (trf == -1) {
        do_res_out(e, x, y, tpop, land_pollution_value(e, x, y));
        return;
    }
    if (t == T_FREEZ || !(rnd16(e) & 7)) {
        int16_t locvalve = eval_res(e, x, y, trf);
        int16_t zscore = (int16_t)(S.resValve + locvalve);
        if (!zonePower) zscore = -500;
        if (zscore > -350 && ((int16_t)(zscore - 26380) > (int16_t)rnd16s(e))) {
            if (!tpop && !(rnd16(e) & 3)) {
                make_hospital(e, x, y);
                return;
            }
            do_res_in(e, x, y, tpop, land_pollution_value(e, x, y));
            return;
        }
        if (zscore < 350 && ((int16_t)(zscore + 26380) < (int16_t)rnd16s(e)))
            do_res_out(e, x, y, tpop, land_pollution_value(e, x, y));
    }
}

// [LEAD] zone.cpp:786-888 doCommercial / doComIn / doComOut
MPN void do_commercial(Env &e, int x, int y, bool zonePower)
{
    MP_COV(do_commercial);
    int t = e.map()[tix(x, y)] & LOMASK;
    S.comZonePop++;
    int16_t tpop = com_zone_pop(t);
    S.comPop = (int16_t)(S.comPop + tpop);
    int16_t trf = (tpop > rnd(e, 5)) ? (int16_t)make_traffic(e, x, y, 1) : (int16_t)1;   // ZT_INDUSTRIAL
    if (trf == -1) {
        int value = land_pollution_value(e, x, y);
        if (tpop > 1) { com_plop(e, x, y, tpop - 2, value); inc_rate_of_growth(e, x, y, -8); }
        else if (tpop == 1) { zone_plop(e, x, y, T_COMBASE); inc_rate_of_growth(e, x, y, -8); }
        return;
    }
    if (!(rnd16(e) & 7)) {
        int16_t locvalve = (trf < 0) ? (int16_t)-3000 : (int16_t)s8get(e.comRate(), x, y);
        int16_t zscore = (int16_t)(S.comValve + locvalve);
        if (!zonePower) zscore = -500;
        if (trf && zscore > -350 && ((int16_t)(zscore - 26380) > (int16_t)rnd16s(e))) {
            int value = land_pollution_value(e, x, y);
            int16_t z = (int16_t)(b2get(e.landval(), x, y) >> 5);
            if (tpop > z) return;
            if (tpop < 5) { com_plop(e, x, y, tpop, value); inc_rate_of_growth(e, x, y, 8); }
            return;
        }
        if (zscore < 350 && ((int16_t)(zscore + 26380) < (int16_t)rnd16s(e))) {
            int value = land_pollution_value(e, x, y);
            if (tpop > 1) { com_plop(e, x, y, tpop - 2, value); inc_rate_of_growth(e, x, y, -8); }
            else if (tpop == 1) { zone_plop(e, x, y, T_COMBASE); inc_rate_of_growth(e, x, y, -8); }
        }
    }
}

// [LEAD] zone.cpp:187-232 setSmoke
MPN void set_smoke(Env &e, int x, int y, bool zonePower)
{
    MP_COV(set_smoke);
    MP_TAB uint8_t aniThis[8] = { 1, 0, 1, 1, 0, 0, 1, 1 };
    MP_TAB int8_t dx1[8] = { -1, 0, 1, 0, 0, 0, 0, 1 }, dy1[8] = { -1, 0, -1, -1, 0, 0, -1, -1 };
    MP_TAB int16_t aniTabB[8] = { 0, 0, 36, 44, 0, 0, 52, 60 };
    MP_TAB int16_t aniTabC[8] = { T_IND1, 0, T_IND2, T_IND4, 0, 0, T_IND6, T_IND8 };
    MP_TAB int16_t aniTabD[8] = { T_IND1, 0, T_IND3, T_IND5, 0, 0, T_IND7, T_IND9 };
    int t = e.map()[tix(x, y)] & LOMASK;
    if (t < T_IZB) return;
    int z = ((t - T_IZB) >> 3) & 7;
    if (!aniThis[z]) return;
    int xx = x + dx1[z], yy = y + dy1[z];
    if (!inb(xx, yy)) return;
    if (zonePower) {
        if ((e.map()[tix(xx, yy)] & LOMASK) == aniTabC[z])
            mset(e, tix(xx, yy), (uint16_t)((ANIMBIT | CONDBIT | BURNBIT) | (T_SMOKEBASE + aniTabB[z])));
    } else {
        if ((e.map()[tix(xx, yy)] & LOMASK) > aniTabC[z])
            mset(e, tix(xx, yy), (uint16_t)((CONDBIT | BURNBIT) | aniTabD[z]));
    }
}

MPD void do_ind_out(Env &e, int x, int y, int pop, int value)      // zone.cpp:969-981
{
    MP_COV(do_ind_out);
    if (pop > 1) { ind_plop(e, x, y, pop - 2, value); inc_rate_of_growth(e, x, y, -8); return; }
    if (pop == 1) { zone_plop(e, x, y, T_INDBASE); inc_rate_of_growth(e, x, y, -8); }
}

// [LEAD] zone.cpp:916-957 doIndustrial
MPN void do_industrial(Env &e, int x, int y, bool zonePower)
{
    MP_COV(do_industrial);
    int t = e.map()[tix(x, y)] & LOMASK;
    S.indZonePop++;
    set_smoke(e, x, y, zonePower);
    int16_t tpop = ind_zone_pop(t);
    S.indPop = (int16_t)(S.indPop + tpop);
    int16_t trf = (tpop > rnd(e, 5)) ? (int16_t)make_traffic(e, x, y, 2) : (int16_t)1;   // ZT_RESIDENTIAL
    if (trf == -1) {
        int v = rnd16(e) & 1;
        do_ind_out(e, x, y, tpop, v);
        return;
    }
    if (!(rnd16(e) & 7)) {
        int16_t zscore = (int16_t)(S.indValve + (trf < 0 ? -1000 : 0));
        if (!zonePower) zscore = -500;
        if (zscore > -350 && ((int16_t)(zscore - 26380) > (int16_t)rnd16s(e))) {
            int v = rnd16(e) & 1;
            if (tpop < 4) { ind_plop(e, x, y, tpop, v); inc_rate_of_growth(e, x, y, 8); }
            return;
        }
        if (zscore < 350 && ((int16_t)(zscore + 26380) < (int16_t)rnd16s(e))) {
            int v = rnd16(e) & 1;
            do_ind_out(e, x, y, tpop, v);
        }
    }
}

// [LEAD] simulate.cpp:1390-1421 repairZone
MPN void repair_zone(Env &e, int x, int y, int zCent, int zSize)
{
    MP_COV(repair_zone);
    int tile = zCent - 2 - zSize;
    MP_NOUNROLL for (int yy = -1; yy < zSize - 1; yy++)
        MP_NOUNROLL for (int xx = -1; xx < zSize - 1; xx++) {
            int ax = x + xx, ay = y + yy;
            tile++;
            if (inb(ax, ay)) {
                int mv = e.map()[tix(ax, ay)];
                if (mv & ZONEBIT) continue;
                if (mv & ANIMBIT) continue;
                int mt = mv & LOMASK;
                if (mt < T_RUBBLE || mt >= T_ROADBASE) mset(e, tix(ax, ay), (uint16_t)(tile | CONDBIT | BURNBIT));
            }
        }
}

// [LEAD] zone.cpp:127-180 doHospitalChurch
MPN void do_hospital_church(Env &e, int x, int y)
{
    MP_COV(do_hospital_church);
    int t = e.map()[tix(x, y)] & LOMASK;
    if (t == T_HOSPITAL) {
        S.hospitalPop++;
        if (!(S.cityTime & 15)) repair_zone(e, x, y, T_HOSPITAL, 3);
        if (S.needHospital == -1 && !rnd(e, 20)) zone_plop(e, x, y, T_RESBASE);
    } else if (t == T_CHURCH0 || t == T_CHURCH1 || t == T_CHURCH2 || t == T_CHURCH3 || t == T_CHURCH4
               || t == T_CHURCH5 || t == T_CHURCH6 || t == T_CHURCH7) {
        S.churchPop++;
        if (!(S.cityTime & 15)) repair_zone(e, x, y, t, 3);
        if (S.needChurch == -1 && !rnd(e, 20)) zone_plop(e, x, y, T_RESBASE);
        // the simulateChurch callback is a no-op for state (micropolisgenericengine.py:1198-1310)
    }
}

// [LEAD] simulate.cpp:1652-1705 doMeltdown
MPN void do_meltdown(Env &e, int x, int y)
{
    MP_COV(do_meltdown);
    make_explosion(e, x - 1, y - 1);
    make_explosion(e, x - 1, y + 2);
    make_explosion(e, x + 2, y - 1);
    make_explosion(e, x + 2, y + 2);
    MP_NOUNROLL for (int xx = x - 1; xx < x + 3; xx++)
        MP_NOUNROLL for (int yy = y - 1; yy < y + 3; yy++) tset(e, xx, yy, random_fire(e));
    MP_NOUNROLL for (int z = 0; z < 200; z++) {
        int xx = x - 20 + rnd(e, 40);
        int yy = y - 15 + rnd(e, 30);
        if (!inb(xx, yy)) continue;
        int t = e.map()[tix(xx, yy)];
        if (t & ZONEBIT) continue;
        if ((t & BURNBIT) || t == T_DIRT) mset(e, tix(xx, yy), T_RADTILE);
    }
}

// [LEAD] simulate.cpp:1430-1585 doSpecialZone (+ drawStadium :1588, doAirport :1606, coalSmoke :1624)
MPN void do_special_zone(Env &e, int x, int y, bool powerOn)
{
    MP_COV(do_special_zone);
    MP_TAB int meltdownTable[3] = { 30000, 20000, 10000 };
    int t = e.map()[tix(x, y)] & LOMASK;
    switch (t) {
    case T_POWERPLANT: {
        S.coalPowerPop++;
        if ((S.cityTime & 7) == 0) repair_zone(e, x, y, T_POWERPLANT, 4);
        push_power_stack(e, x, y);
        MP_TAB int16_t sm[4] = { T_COALSMOKE1, T_COALSMOKE2, T_COALSMOKE3, T_COALSMOKE4 };
        MP_TAB int8_t dx[4] = { 1, 2, 1, 2 }, dy[4] = { -1, -1, 0, 0 };
        MP_NOUNROLL for (int i = 0; i < 4; i++)
            tset(e, x + dx[i], y + dy[i], sm[i] | ANIMBIT | CONDBIT | PWRBIT | BURNBIT);
        return;
    }
    case T_NUCLEAR:
        if (S.enableDisasters && !rnd(e, (int16_t)meltdownTable[S.gameLevel])) {
            do_meltdown(e, x, y);
            return;
        }
        S.nuclearPowerPop++;
        if ((S.cityTime & 7) == 0) repair_zone(e, x, y, T_NUCLEAR, 4);
        push_power_stack(e, x, y);
        return;
    case T_FIRESTATION:
    case T_POLICESTATION: {
        bool fire = (t == T_FIRESTATION);
        if (fire) S.fireStationPop++; else S.policeStationPop++;
        if (!(S.cityTime & 7)) repair_zone(e, x, y, t, 3);
        int z = (int)(fire ? S.fireEffect : S.policeEffect);
        if (!powerOn) z = z / 2;
        int px = x, py = y;
        if (!find_perimeter_road(e, px, py)) z = z / 2;
        int16_t *m = fire ? e.fireSt() : e.policeSt();
        s8set(m, px, py, s8get(m, px, py) + z);
        return;
    }
    case T_STADIUM:
        S.stadiumPop++;
        if (!(S.cityTime & 15)) repair_zone(e, x, y, T_STADIUM, 4);
        if (powerOn && ((S.cityTime + x + y) & 31) == 0) {
            int z = T_FULLSTADIUM - 5;
            MP_NOUNROLL for (int yy = y - 1; yy < y + 3; yy++)
                MP_NOUNROLL for (int xx = x - 1; xx < x + 3; xx++) { tset(e, xx, yy, z | BNCNBIT); z++; }
            mset(e, tix(x, y), e.map()[tix(x, y)] | ZONEBIT | PWRBIT);
            tset(e, x + 1, y, T_FOOTBALLGAME1 + ANIMBIT);
            tset(e, x + 1, y + 1, T_FOOTBALLGAME2 + ANIMBIT);
        }
        return;
    case T_FULLSTADIUM:
        S.stadiumPop++;
        if (((S.cityTime + x + y) & 7) == 0) {
            int z = T_STADIUM - 5;
            MP_NOUNROLL for (int yy = y - 1; yy < y + 3; yy++)
                MP_NOUNROLL for (int xx = x - 1; xx < x + 3; xx++) { tset(e, xx, yy, z | BNCNBIT); z++; }
            mset(e, tix(x, y), e.map()[tix(x, y)] | ZONEBIT | PWRBIT);
        }
        return;
    case T_AIRPORT:
        S.airportPop++;
        if ((S.cityTime & 7) == 0) repair_zone(e, x, y, T_AIRPORT, 6);
        if (powerOn) {
            if ((tget(e, x + 1, y - 1) & LOMASK) == T_RADAR)
                tset(e, x + 1, y - 1, T_RADAR0 + ANIMBIT + CONDBIT + BURNBIT);
        } else {
            tset(e, x + 1, y - 1, T_RADAR + CONDBIT + BURNBIT);
        }
        if (powerOn) {
            if (rnd(e, 5) == 0) { generate_plane(e, x, y); return; }
            if (rnd(e, 12) == 0) generate_copter(e, x, y);
        }
        return;
    case T_PORT:
        S.seaportPop++;
        if ((S.cityTime & 15) == 0) repair_zone(e, x, y, T_PORT, 4);
        if (powerOn && get_sprite(e, SPR_SHIP) < 0) generate_ship(e);
        return;
    default:
        return;
    }
}

// [LEAD] zone.cpp:79-120 doZone
MPQ void do_zone(Env &e, int x, int y)
{
    MP_COV(do_zone);
    bool pw = set_zone_power(e, x, y);
    if (pw) S.poweredZoneCount++; else S.unpoweredZoneCount++;
    int t = e.map()[tix(x, y)] & LOMASK;
    if (t > T_PORTBASE && t < T_CHURCH1BASE) { do_special_zone(e, x, y, pw); return; }
    if (t < T_HOSPITAL) { do_residential(e, x, y, pw); return; }
    if (t < T_COMBASE || (t >= T_CHURCH1BASE && t <= T_CHURCH7LAST)) { do_hospital_church(e, x, y); return; }
    if (t < T_INDBASE) { do_commercial(e, x, y, pw); return; }
    if (t < T_CHURCH1BASE) { do_industrial(e, x, y, pw); return; }
}

// ------------------------------------------------------------------ fire / flood / radiation
// [LEAD] simulate.cpp:1336-1381 fireZone (also disasters.cpp doFlood)
MPN void fire_zone(Env &e, int x, int y, int ch)
{
    MP_COV(fire_zone);
    int v = tclamp(s8get(e.rog(), x, y) - 20, -200, 200);
    s8set(e.rog(), x, y, v);
    ch &= LOMASK;
    int xymax = ch < T_PORTBASE ? 2 : (ch == T_AIRPORT ? 5 : 4);
    MP_NOUNROLL for (int dx = -1; dx < xymax; dx++)
        MP_NOUNROLL for (int dy = -1; dy < xymax; dy++) {
            int xx = x + dx, yy = y + dy;
            if (inb(xx, yy) && (e.map()[tix(xx, yy)] & LOMASK) >= T_ROADBASE) e.map()[tix(xx, yy)] |= BULLBIT;
        }
}

// [LEAD] simulate.cpp:1280-1327 doFire
MPN void do_fire(Env &e, int x, int y)
{
    MP_COV(do_fire);
    MP_TAB int8_t dx[4] = { -1, 0, 1, 0 }, dy[4] = { 0, -1, 0, 1 };
    MP_NOUNROLL for (int z = 0; z < 4; z++) {
        if ((rnd16(e) & 7) == 0) {
            int xx = x + dx[z], yy = y + dy[z];
            if (inb(xx, yy)) {
                int c = e.map()[tix(xx, yy)];
                if (!(c & BURNBIT)) continue;
                if (c & ZONEBIT) {
                    fire_zone(e, xx, yy, c);
                    if ((c & LOMASK) > T_IZB) make_explosion_at(e, xx * 16 + 8, yy * 16 + 8);
                }
                mset(e, tix(xx, yy), (uint16_t)random_fire(e));
            }
        }
    }
    // getRandom(rate), rate = 10 / 3 / 2 / 1 by fire-station coverage: a literal range per branch (a runtime range
    // costs a 16-bit division and a remainder subroutine call)
    const int z = s8get(e.fireEff(), x, y);
    int r;
    if (z > 100) r = rnd(e, 1);
    else if (z > 20) r = rnd(e, 2);
    else if (z > 0) r = rnd(e, 3);
    else r = rnd(e, 10);
    if (r == 0) mset(e, tix(x, y), (uint16_t)random_rubble(e));
}

// [LEAD] disasters.cpp:375-405 doFlood
MPN void do_flood(Env &e, int x, int y)
{
    MP_COV(do_flood);
    MP_TAB int8_t dx[4] = { 0, 1, 0, -1 }, dy[4] = { -1, 0, 1, 0 };
    if (S.floodCount > 0) {
        MP_NOUNROLL for (int z = 0; z < 4; z++) {
            if ((rnd16(e) & 7) == 0) {
                int xx = x + dx[z], yy = y + dy[z];
                if (inb(xx, yy)) {
                    int c = e.map()[tix(xx, yy)], t = c & LOMASK;
                    if ((c & BURNBIT) == BURNBIT || c == T_DIRT || (t >= T_WOODS5 && t < T_FLOOD)) {
                        if ((c & ZONEBIT) == ZONEBIT) fire_zone(e, xx, yy, c);
                        mset(e, tix(xx, yy), (uint16_t)(T_FLOOD + rnd(e, 2)));
                    }
                }
            }
        }
    } else if ((rnd16(e) & 15) == 0) {
        mset(e, tix(x, y), T_DIRT);
    }
}

// ------------------------------------------------------------------ roads / rails
// [LEAD] simulate.cpp:1114-1245 doBridge
MPN bool do_bridge(Env &e, int x, int y, int tile)
{
    MP_COV(do_bridge);
    MP_TAB int8_t HDx[7] = { -2, 2, -2, -1, 0, 1, 2 }, HDy[7] = { -1, -1, 0, 0, 0, 0, 0 };
    MP_TAB int16_t HBRTAB[7] = { T_HBRDG1 | BULLBIT, T_HBRDG3 | BULLBIT, T_HBRDG0 | BULLBIT, T_RIVER,
        T_BRWH | BULLBIT, T_RIVER, T_HBRDG2 | BULLBIT };
    MP_TAB int16_t HBRTAB2[7] = { T_RIVER, T_RIVER, T_HBRIDGE | BULLBIT, T_HBRIDGE | BULLBIT,
        T_HBRIDGE | BULLBIT, T_HBRIDGE | BULLBIT, T_HBRIDGE | BULLBIT };
    MP_TAB int8_t VDx[7] = { 0, 1, 0, 0, 0, 0, 1 }, VDy[7] = { -2, -2, -1, 0, 1, 2, 2 };
    MP_TAB int16_t VBRTAB[7] = { T_VBRDG0 | BULLBIT, T_VBRDG1 | BULLBIT, T_RIVER, T_BRWV | BULLBIT,
        T_RIVER, T_VBRDG2 | BULLBIT, T_VBRDG3 | BULLBIT };
    MP_TAB int16_t VBRTAB2[7] = { T_VBRIDGE | BULLBIT, T_RIVER, T_VBRIDGE | BULLBIT, T_VBRIDGE | BULLBIT,
        T_VBRIDGE | BULLBIT, T_VBRIDGE | BULLBIT, T_RIVER };
    if (tile == T_BRWV) {
        if (!(rnd16(e) & 3) && boat_distance(e, x, y) > 340)
            MP_NOUNROLL for (int z = 0; z < 7; z++) {
                int xx = x + VDx[z], yy = y + VDy[z];
                if (inb(xx, yy) && (e.map()[tix(xx, yy)] & LOMASK) == (VBRTAB[z] & LOMASK))
                    mset(e, tix(xx, yy), (uint16_t)VBRTAB2[z]);
            }
        return true;
    }
    if (tile == T_BRWH) {
        if (!(rnd16(e) & 3) && boat_distance(e, x, y) > 340)
            MP_NOUNROLL for (int z = 0; z < 7; z++) {
                int xx = x + HDx[z], yy = y + HDy[z];
                if (inb(xx, yy) && (e.map()[tix(xx, yy)] & LOMASK) == (HBRTAB[z] & LOMASK))
                    mset(e, tix(xx, yy), (uint16_t)HBRTAB2[z]);
            }
        return true;
    }
    if (boat_distance(e, x, y) < 300 || !(rnd16(e) & 7)) {
        if (tile & 1) {
            if (x < WORLD_W - 1 && e.map()[tix(x + 1, y)] == T_CHANNEL) {
                MP_NOUNROLL for (int z = 0; z < 7; z++) {
                    int xx = x + VDx[z], yy = y + VDy[z];
                    if (inb(xx, yy)) {
                        int m = e.map()[tix(xx, yy)];
                        if (m == T_CHANNEL || ((m & 15) == (VBRTAB2[z] & 15))) mset(e, tix(xx, yy), (uint16_t)VBRTAB[z]);
                    }
                }
                return true;
            }
            return false;
        } else {
            if (y > 0 && e.map()[tix(x, y - 1)] == T_CHANNEL) {
                MP_NOUNROLL for (int z = 0; z < 7; z++) {
                    int xx = x + HDx[z], yy = y + HDy[z];
                    if (inb(xx, yy)) {
                        int m = e.map()[tix(xx, yy)];
                        if (((m & 15) == (HBRTAB2[z] & 15)) || m == T_CHANNEL) mset(e, tix(xx, yy), (uint16_t)HBRTAB[z]);
                    }
                }
                return true;
            }
            return false;
        }
    }
    return false;
}

// [LEAD] simulate.cpp:1048-1111 doRoad
MPQ void do_road(Env &e, int x, int y)
{
    MP_COV(do_road);
    MP_TAB int16_t densityTable[3] = { T_ROADBASE, T_LTRFBASE, T_HTRFBASE };
    S.roadTotal++;
    int mv = e.map()[tix(x, y)], tile = mv & LOMASK;
    if (S.roadEffect < (15 * MAX_ROAD_EFFECT / 16)) {
        if ((rnd16(e) & 511) == 0) {
            if (!(mv & CONDBIT)) {
                if (S.roadEffect < (rnd16(e) & 31)) {
                    if ((tile & 15) < 2 || (tile & 15) == 15) mset(e, tix(x, y), T_RIVER);
                    else mset(e, tix(x, y), (uint16_t)random_rubble(e));
                    return;
                }
            }
        }
    }
    if ((mv & BURNBIT) == 0) {
        S.roadTotal = (int16_t)(S.roadTotal + 4);
        if (do_bridge(e, x, y, tile)) return;
    }
    int16_t tden;
    if (tile < T_LTRFBASE) tden = 0;
    else if (tile < T_HTRFBASE) tden = 1;
    else { S.roadTotal++; tden = 2; }
    int16_t td = (int16_t)(b2get(e.traffic(), x, y) >> 6);
    if (td > 1) td--;
    if (tden != td) {
        int16_t z = (int16_t)(((tile - T_ROADBASE) & 15) + densityTable[td]);
        z |= mv & (ALLBITS - ANIMBIT);
        if (td > 0) z |= ANIMBIT;
        mset(e, tix(x, y), (uint16_t)z);
    }
}

// sprite.cpp:1831-1837 generateTrain: the guard is inline (every rail tile of every scan asks), the sprite is made out of line
MPD void generate_train(Env &e, int x, int y)
{
    MP_COV(generate_train);
    const int ti = S.globalSprites[SPR_TRAIN];                  // getSprite (sprite.cpp:336-344)
    if (S.totalPop > 20 && (ti < 0 || e.spr()[ti].frame == 0) && rnd(e, 25) == 0) generate_train_make(e, x, y);
}

// [LEAD] simulate.cpp:1005-1037 doRail
MPQ void do_rail(Env &e, int x, int y)
{
    MP_COV(do_rail);
    S.railTotal++;
    generate_train(e, x, y);
    if (S.roadEffect < (15 * MAX_ROAD_EFFECT / 16)) {
        if (!(rnd16(e) & 511)) {
            int cv = e.map()[tix(x, y)];
            if (!(cv & CONDBIT)) {
                if (S.roadEffect < (rnd16(e) & 31)) {
                    int t = cv & LOMASK;
                    if (t < T_RAILBASE + 2) mset(e, tix(x, y), T_RIVER);
                    else mset(e, tix(x, y), (uint16_t)random_rubble(e));
                }
            }
        }
    }
}

// ------------------------------------------------------------------ map scan

// [LEAD] simulate.cpp:929-997 body of mapScan for one tile
MPP void scan_tile(Env &e, int idx)
{
    MP_COV(scan_tile);
    int mv = e.map()[idx], tile = mv & LOMASK;
    int x = idx / WORLD_H, y = idx - x * WORLD_H;
    if (tile < T_ROADBASE) {
        if (tile >= T_FIREBASE) {
            S.firePop++;
            if (!(rnd16(e) & 3)) do_fire(e, x, y);
            return;
        }
        if (tile < T_RADTILE) {
            // doFlood (disasters.cpp:375-405) without the call when nothing happens: a receding flood
            // tile is one draw (1 in 16 dries up), an advancing one is four draws of which each spreads
            // 1 in 8 -- 59 % of the visits spread nowhere, and then the four draws are all there is
            if (S.floodCount > 0) {
                uint32_t r = S.rng;
                bool hit = false;
                MP_NOUNROLL for (int z = 0; z < 4; z++) {
                    r = r * 1103515245u + 12345u;
                    hit |= (r & 0x700u) == 0;                          // (draw >> 8) & 7 == 0
                }
                if (hit) do_flood(e, x, y);                            // replays the draws from S.rng
                else S.rng = r;
            } else if ((rnd16(e) & 15) == 0) {
                mset(e, idx, T_DIRT);
            }
        } else if ((rnd16(e) & 4095) == 0) {
            mset(e, idx, T_DIRT);                                      // doRadTile :1040-1045
        }
        return;
    }
    // (a zone centre gets its setZonePower from doZone right below: the call is idempotent and draws nothing)
    if (S.newPower && (mv & CONDBIT) && (!(mv & ZONEBIT) || tile < T_POWERBASE)) set_zone_power(e, x, y);
    if (tile >= T_ROADBASE && tile < T_POWERBASE) { do_road(e, x, y); return; }
    if (mv & ZONEBIT) { do_zone(e, x, y); return; }
    if (tile >= T_RAILBASE && tile < T_RESBASE) { do_rail(e, x, y); return; }
    if (tile >= T_SOMETINYEXP && tile <= T_LASTTINYEXP) { mset(e, idx, (uint16_t)random_rubble(e)); return; }
    // a building tile that is no centre: setZonePower above was all there is to do, and it is done
    e.need()[idx >> 5] &= ~(1u << (idx & 31));
}

// [ALL] mapScan(x1, x2).  The scan order is linear memory order (x outer, y inner over the
// column-major map), and the reference skips every tile whose id is below FLOOD
// (simulate.cpp:937-945) -- on a 16x16 build window that is ~98% of the 12 000 tiles.  The active-tile
// bitmap (Env::act, maintained by mset) turns the scan into "find the next set bit": each lane
// loads one 32-tile word of the range, a ballot finds the first word with work, lane 0 processes
// that tile, and the words are re-read (L1 hits) because processing may have activated tiles ahead
// of the scan position (fire spread, zonePlop, repairZone ...).  Tiles activated BEHIND the scan
// position are not visited, exactly as in the reference's single forward sweep.
#ifndef MP_DENSE_RANGE_MIN
#define MP_DENSE_RANGE_MIN 96       // tiles needing a visit in one scan range before the batching scan is used
#endif
#ifndef MP_DENSE_WORD_MIN
#define MP_DENSE_WORD_MIN 4         // tiles needing a visit in one 32-tile word before the word is batched
#endif

// ---- dense scan -------------------------------------------------------------------------------
// A flood, a meltdown or a forest fire turns hundreds to thousands of tiles of one city active; on
// the serial lane such a city costs 10x the average and the whole batch waits for it at the end of
// every step (tools/imbalance.py: max/mean = 10).  Most of those tiles are cheap and touch nothing
// but themselves, so when a scan range holds many tiles that need a visit the words are processed
// 32 tiles at a time: lane b takes tile b, the draws are handed out in scan order by LCG
// jump-ahead, and the batch stops in front of the first tile that really needs the serial lane:
//   * building tile that is no centre: at most a setZonePower          (no draw)
//   * plain road: census + traffic art                                  (no draw)
//   * receding flood (floodCount == 0), radioactive waste, tiny explosion, fire that does not
//     spread this time                                                   (1 draw each)
//   * ACTIVE flooding (floodCount > 0): 4 draws; the tile is a no-op unless one of them hits
//     (1 in 8) on a neighbour that can still be flooded -- in the interior of a flood field no
//     neighbour can, so whole words of flood water are retired at once   (4 draws)
//   * everything else (zones, bridges, rails, spreading fire / flood): serial lane.
enum : int { TK_SERIAL = 0, TK_INERT, TK_ROAD, TK_FLOOD0, TK_RAD, TK_FIRE, TK_TINYEXP, TK_FLOODA };

MPD int tile_kind(const Env &e, int v)
{
    MP_COV(tile_kind);
    const int t = v & LOMASK;
    const bool flooding = S.floodCount > 0;
    if (t < T_ROADBASE) {
        if (t >= T_FIREBASE) return TK_FIRE;
        if (t < T_RADTILE) return flooding ? (int)TK_FLOODA : (int)TK_FLOOD0;
        return flooding ? (int)TK_SERIAL : (int)TK_RAD;          // a tile that turns to DIRT could be flooded by a later lane
    }
    if (t < T_POWERBASE) {
        if (S.roadEffect < (15 * MAX_ROAD_EFFECT / 16) || !(v & BURNBIT)) return TK_SERIAL;   // decay draw / bridge
        return TK_ROAD;
    }
    if (v & ZONEBIT) return TK_SERIAL;
    if (t >= T_RAILBASE && t < T_RESBASE) return TK_SERIAL;     // doRail: conditional draw
    if (t >= T_SOMETINYEXP && t <= T_LASTTINYEXP) return flooding ? (int)TK_SERIAL : (int)TK_TINYEXP;
    return TK_INERT;
}

#if MP_WARP_COOP
// n steps of the engine's LCG (random.cpp:88-100) in O(log n): step^(2^j) is the affine map
// x -> A_j x + C_j with A_{j+1} = A_j^2, C_{j+1} = (A_j + 1) C_j; n < 256.
struct LcgPow { uint32_t a[8], c[8]; };
__host__ __device__ constexpr LcgPow lcg_pow_table()
{
    LcgPow t = {};
    uint32_t a = 1103515245u, c = 12345u;
    for (int j = 0; j < 8; j++) {
        t.a[j] = a; t.c[j] = c;
        c = (a + 1u) * c;
        a = a * a;
    }
    return t;
}
MPD uint32_t lcg_jump(uint32_t x, int n)
{
    MP_COV(lcg_jump);
    constexpr LcgPow t = lcg_pow_table();
#pragma unroll
    for (int j = 0; j < 8; j++)
        if ((n >> j) & 1) x = x * t.a[j] + t.c[j];
    return x;
}
MPD int lcg_r16(uint32_t x) { return (int)((x & 0xffff00u) >> 8) & 0xffff; }

// [ALL lanes] one 32-tile word: lane b takes tile b of word `wi` (bits of mL, already limited to the
// unscanned part of the range).  Returns the scan position after the batch; serial_next says whether
// the tile at that position is the one the batch had to stop in front of.
MPD int scan_word_batch(Env &e, int wi, uint32_t mL, bool &serial_next)
{
    MP_COV(scan_word_batch);
    const int lane = e.c.tid;
    const bool mine = (mL >> lane) & 1u;
    const int t = wi * 32 + lane;
    int v = mine ? e.map()[t] : 0;
    const int kind = mine ? tile_kind(e, v) : (int)TK_INERT;
    const unsigned serial = __ballot_sync(0xffffffffu, mine && kind == TK_SERIAL);
    int stop = serial ? __ffs(serial) - 1 : 32;             // batch = my tiles below `stop`
    // draws per tile and their rank in scan order (exclusive prefix sum over the lanes)
    const int nd = (mine && lane < stop) ? (kind == TK_FLOODA ? 4 : (kind >= TK_FLOOD0 ? 1 : 0)) : 0;
    int pre = nd;
    MP_NOUNROLL for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += y;
    }
    pre -= nd;
    int r = 0;
    bool must_serial = false;
    if (nd) {
        uint32_t x = lcg_jump(S.rng, pre + 1);
        r = lcg_r16(x);
        if (kind == TK_FIRE) {
            must_serial = !(r & 3);                         // 1 in 4: doFire, which draws on
        } else if (kind == TK_FLOODA) {                     // doFlood, floodCount > 0 (disasters.cpp:382-400)
            const int tx = t / WORLD_H, ty = t - tx * WORLD_H;
            MP_NOUNROLL for (int z = 0; z < 4; z++) {
                if ((r & 7) == 0) {
                    const int xx = tx + (z == 1) - (z == 3), yy = ty - (z == 0) + (z == 2);
                    if (inb(xx, yy)) {
                        const int c = e.map()[tix(xx, yy)], ct = c & LOMASK;
                        if ((c & BURNBIT) == BURNBIT || c == T_DIRT || (ct >= T_WOODS5 && ct < T_FLOOD)) must_serial = true;
                    }
                }
                x = x * 1103515245u + 12345u;
                r = lcg_r16(x);
            }
        }
    }
    const unsigned ms = __ballot_sync(0xffffffffu, must_serial);
    if (ms) stop = min(stop, __ffs(ms) - 1);
    const bool batch = mine && lane < stop;
    bool still = true, heavy = false;
    if (batch) {
        if (kind <= TK_ROAD) {
            if (S.newPower && (v & CONDBIT)) {              // setZonePower (zone.cpp:419-437)
                const int id = v & LOMASK;
                const bool on = id == T_NUCLEAR || id == T_POWERPLANT || ((e.power()[t >> 5] >> (t & 31)) & 1u);
                const int nv = on ? (v | PWRBIT) : (v & ~PWRBIT);
                if (nv != v) { v = nv; e.map()[t] = (uint16_t)v; }
            }
            if (kind == TK_ROAD) {                          // doRoad (simulate.cpp:1048-1111)
                const int tile = v & LOMASK;
                int tden = tile < T_LTRFBASE ? 0 : (tile < T_HTRFBASE ? 1 : 2);
                heavy = tden == 2;
                const int tx = t / WORLD_H, ty = t - tx * WORLD_H;
                int td = e.traffic()[(tx >> 1) * H2 + (ty >> 1)] >> 6;
                if (td > 1) td--;
                if (tden != td) {
                    int z = ((tile - T_ROADBASE) & 15) + (td == 0 ? (int)T_ROADBASE : (td == 1 ? (int)T_LTRFBASE : (int)T_HTRFBASE));
                    z |= v & (ALLBITS - ANIMBIT);
                    if (td > 0) z |= ANIMBIT;
                    e.map()[t] = (uint16_t)z;
                }
            }
        } else if (kind == TK_FLOOD0) {                     // doFlood, floodCount == 0
            if ((r & 15) == 0) { e.map()[t] = T_DIRT; still = false; }
        } else if (kind == TK_RAD) {                        // doRadTile
            if ((r & 4095) == 0) { e.map()[t] = T_DIRT; still = false; }
        } else if (kind == TK_TINYEXP) {                    // tiny explosion -> rubble
            e.map()[t] = (uint16_t)((T_RUBBLE + (r & 3)) | BULLBIT);
            still = false;
        }                                                   // TK_FIRE (no spread): firePop++; TK_FLOODA: no-op
    }
    const unsigned gone = __ballot_sync(0xffffffffu, batch && !still);
    const unsigned inert_done = __ballot_sync(0xffffffffu, batch && kind == TK_INERT);
    const int roads = __popc(__ballot_sync(0xffffffffu, batch && kind == TK_ROAD))
        + __popc(__ballot_sync(0xffffffffu, batch && heavy));
    const int fires = __popc(__ballot_sync(0xffffffffu, batch && kind == TK_FIRE));
    // draws consumed by the batch = prefix of the lane it stopped at (or the word total)
    const int total = __shfl_sync(0xffffffffu, pre + nd, 31);
    const int used = stop < 32 ? __shfl_sync(0xffffffffu, pre, stop) : total;
    if (lane == 0) {
        if (gone) e.act()[wi] &= ~gone;
        if (gone | inert_done) e.need()[wi] &= ~(gone | inert_done);
        if (roads) S.roadTotal = (int16_t)(S.roadTotal + roads);
        if (fires) S.firePop = (int16_t)(S.firePop + fires);
        S.rng = lcg_jump(S.rng, used);
    }
    __syncwarp();
    serial_next = stop < 32 && ((mL >> stop) & 1u);
    return serial_next ? wi * 32 + stop : (wi + 1) * 32;
}

// [ALL] the whole mapScan range with word batching (see above).  Out of line and entered only when
// the range is dense, so the common sparse loop below carries none of this.
MPN void map_scan_dense(Env &e, int i0, int i1)
{
    MP_COV(map_scan_dense);
    const int w0 = i0 >> 5, w1 = (i1 - 1) >> 5;
    const int lane = e.c.tid;
    int pos = i0;
    MP_NOUNROLL for (int wb = w0; wb <= w1; wb += 32) {
        const int w = wb + lane;
        for (;;) {
            uint32_t m = (w <= w1) ? e.need()[w] : 0u;
            const int lo = pos - w * 32, hi = i1 - w * 32;
            if (lo > 0) m &= (lo >= 32) ? 0u : (0xffffffffu << lo);
            if (hi < 32) m &= (hi <= 0) ? 0u : (0xffffffffu >> (32 - hi));
            const unsigned any = __ballot_sync(0xffffffffu, m != 0);
            if (!any) break;
            const int L = __ffs(any) - 1;
            const uint32_t mL = __shfl_sync(0xffffffffu, m, L);
            int t1 = (wb + L) * 32 + (__ffs(mL) - 1);
            if (__popc(mL) >= MP_DENSE_WORD_MIN) {
                bool serial_next;
                t1 = scan_word_batch(e, wb + L, mL, serial_next);
                if (!serial_next) { pos = t1; continue; }
            }
            if (lane == 0) scan_tile(e, t1);
            __syncwarp();
            pos = t1 + 1;
        }
    }
}
#endif

MPP void map_scan(Env &e, int x1, int x2)
{
    MP_COV(map_scan);
    const int i0 = x1 * WORLD_H, i1 = x2 * WORLD_H;
    const int w0 = i0 >> 5, w1 = (i1 - 1) >> 5;
    int pos = i0;                                   // first tile not yet passed by the scan
#if MP_WARP_COOP
    const int lane = e.c.tid;
    {   // one look at the strip's words (<= 64): nothing to visit (the usual case outside the build
        // window) ends the phase here; a dense range (flood / radiation / fire field) goes to the
        // batching scan
        const int wa = w0 + lane, wb2 = w0 + 32 + lane;
        uint32_t a = (wa <= w1) ? e.need()[wa] : 0u, b = (wb2 <= w1) ? e.need()[wb2] : 0u;
        if (wa == w0) a &= 0xffffffffu << (i0 & 31);
        if (wa == w1 && (i1 & 31)) a &= 0xffffffffu >> (32 - (i1 & 31));
        if (wb2 == w1 && (i1 & 31)) b &= 0xffffffffu >> (32 - (i1 & 31));
        if (!__any_sync(0xffffffffu, (a | b) != 0)) return;
        int cnt = __popc(a) + __popc(b);
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (cnt >= MP_DENSE_RANGE_MIN) {
            map_scan_dense(e, i0, i1);
            e.c.sync();
            return;
        }
    }
    MP_NOUNROLL for (int wb = w0; wb <= w1; wb += 32) {
        const int w = wb + lane;
        // tiles in [i0, i1): only the first and the last word of the strip are cut, and the scan position moves by
        // whole words (the serial lane finishes the word it takes), so the per-iteration mask is "words from wn on"
        uint32_t edge = 0xffffffffu;
        if (w == w0) edge &= 0xffffffffu << (i0 & 31);
        if (w == w1 && (i1 & 31)) edge &= 0xffffffffu >> (32 - (i1 & 31));
        if (w > w1) edge = 0u;
        int wn = pos >> 5;                                  // first word not yet passed (pos is i0 or a word boundary)
        for (;;) {
            const uint32_t m = (w >= wn && edge) ? (e.need()[w] & edge) : 0u;
            const unsigned any = __ballot_sync(0xffffffffu, m != 0);
            if (!any) break;
            const int wi = wb + __ffs(any) - 1;             // first word with work: the serial lane takes all of it
            if (lane == 0) {
                const int hiw = i1 - wi * 32;
                const uint32_t top = hiw < 32 ? (0xffffffffu >> (32 - hiw)) : 0xffffffffu;
                int p = pos - wi * 32;                      // bits below p are behind the scan
                if (p < 0) p = 0;
                for (;;) {
                    // re-read: the visit may have activated tiles ahead of the scan position
                    const uint32_t mm = e.need()[wi] & top & (0xffffffffu << p);
                    if (!mm) break;
                    const int bit = __ffs(mm) - 1;
                    scan_tile(e, wi * 32 + bit);
                    if (bit == 31) break;
                    p = bit + 1;
                }
            }
            __syncwarp();
            pos = (wi + 1) * 32;
            wn = wi + 1;
        }
    }
#else
    for (int w = w0; w <= w1; w++) {
        for (;;) {
            uint32_t m = e.need()[w];
            const int lo = pos - w * 32, hi = i1 - w * 32;
            if (lo > 0) m &= (lo >= 32) ? 0u : (0xffffffffu << lo);
            if (hi < 32) m &= (hi <= 0) ? 0u : (0xffffffffu >> (32 - hi));
            if (!m) break;
#if defined(__CUDA_ARCH__)
            const int b = __ffs(m) - 1;
#else
            int b = 0;
            while (!((m >> b) & 1u)) b++;
#endif
            const int t = w * 32 + b;
            scan_tile(e, t);
            pos = t + 1;
        }
    }
#endif
    e.c.sync();
}

// ------------------------------------------------------------------ phase 9: census, tax
// [ALL] history scroll shared by take10Census / take120Census: hist[x + 1] = hist[x] for x = hi..lo
// on all six graphs at once (read everything, barrier, write).
MPD void scroll_history(Env &e, int lo, int hi)
{
    MP_COV(scroll_history);
    int16_t *h[6] = { e.resHist(), e.comHist(), e.indHist(), e.crimeHist(), e.pollutionHist(), e.moneyHist() };
    const int n = hi - lo + 1;                       // <= 119 entries per graph: 4 per lane
    int16_t v[6][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int x = lo + e.c.tid + k * e.c.n;
#pragma unroll
        for (int g = 0; g < 6; g++) v[g][k] = (x <= hi) ? h[g][x] : (int16_t)0;
    }
    e.c.sync();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int x = lo + e.c.tid + k * e.c.n;
        if (x <= hi) {
#pragma unroll
            for (int g = 0; g < 6; g++) h[g][x + 1] = v[g][k];
        }
    }
    (void)n;
    e.c.sync();
}

// [ALL] simulate.cpp:712-780 take10Census
MPN void take10_census(Env &e)
{
    MP_COV(take10_census);
#if MP_WARP_COOP
    scroll_history(e, 0, 118);
#else
    for (int x = 118; x >= 0; x--) {
        e.resHist()[x + 1] = e.resHist()[x]; e.comHist()[x + 1] = e.comHist()[x];
        e.indHist()[x + 1] = e.indHist()[x]; e.crimeHist()[x + 1] = e.crimeHist()[x];
        e.pollutionHist()[x + 1] = e.pollutionHist()[x]; e.moneyHist()[x + 1] = e.moneyHist()[x];
    }
#endif
    if (!e.c.lead()) return;
    e.resHist()[0] = (int16_t)(S.resPop / 8);
    e.comHist()[0] = S.comPop;
    e.indHist()[0] = S.indPop;
    S.crimeRamp = (int16_t)(S.crimeRamp + (S.crimeAverage - S.crimeRamp) / 4);
    e.crimeHist()[0] = tmin(S.crimeRamp, (int16_t)255);
    S.pollutionRamp = (int16_t)(S.pollutionRamp + (S.pollutionAverage - S.pollutionRamp) / 4);
    e.pollutionHist()[0] = tmin(S.pollutionRamp, (int16_t)255);
    int16_t x = (int16_t)((S.cashFlow / 20) + 128);
    e.moneyHist()[0] = tclamp(x, (int16_t)0, (int16_t)255);
    int16_t resPopScaled = (int16_t)(S.resPop >> 8);
    if (S.hospitalPop < resPopScaled) S.needHospital = 1;
    if (S.hospitalPop > resPopScaled) S.needHospital = -1;
    if (S.hospitalPop == resPopScaled) S.needHospital = 0;
    int faithfulPop = resPopScaled + S.faith;
    if (S.churchPop < faithfulPop) S.needChurch = 1;
    if (S.churchPop > faithfulPop) S.needChurch = -1;
    if (S.churchPop == faithfulPop) S.needChurch = 0;
}

// [ALL] simulate.cpp:784-826 take120Census
MPN void take120_census(Env &e)
{
    MP_COV(take120_census);
#if MP_WARP_COOP
    scroll_history(e, 120, 238);
#else
    for (int x = 238; x >= 120; x--) {
        e.resHist()[x + 1] = e.resHist()[x]; e.comHist()[x + 1] = e.comHist()[x];
        e.indHist()[x + 1] = e.indHist()[x]; e.crimeHist()[x + 1] = e.crimeHist()[x];
        e.pollutionHist()[x + 1] = e.pollutionHist()[x]; e.moneyHist()[x + 1] = e.moneyHist()[x];
    }
#endif
    if (!e.c.lead()) return;
    e.resHist()[120] = (int16_t)(S.resPop / 8);
    e.comHist()[120] = S.comPop;
    e.indHist()[120] = S.indPop;
    e.crimeHist()[120] = e.crimeHist()[0];
    e.pollutionHist()[120] = e.pollutionHist()[0];
    e.moneyHist()[120] = e.moneyHist()[0];
}

// [LEAD] budget.cpp:107-316 doBudgetNow(false)
MPN void do_budget(Env &e)
{
    MP_COV(do_budget);
    int64_t fireInt = (int)(S.fireFund * S.firePercent);
    int64_t policeInt = (int)(S.policeFund * S.policePercent);
    int64_t roadInt = (int)(S.roadFund * S.roadPercent);
    int64_t total = fireInt + policeInt + roadInt;
    int64_t yum = S.taxFund + S.totalFunds;
    if (yum > total) {
        S.fireValue = fireInt; S.policeValue = policeInt; S.roadValue = roadInt;
    } else if (total > 0) {
        if (yum > roadInt) {
            S.roadValue = roadInt;
            yum -= roadInt;
            if (yum > fireInt) {
                S.fireValue = fireInt;
                yum -= fireInt;
                if (yum > policeInt) {
                    S.policeValue = policeInt;
                    yum -= policeInt;
                } else {
                    S.policeValue = yum;
                    S.policePercent = yum > 0 ? ((float)yum) / ((float)S.policeFund) : 0.0f;
                }
            } else {
                S.fireValue = yum;
                S.policeValue = 0;
                S.policePercent = 0.0f;
                S.firePercent = yum > 0 ? ((float)yum) / ((float)S.fireFund) : 0.0f;
            }
        } else {
            S.roadValue = yum;
            S.fireValue = 0; S.policeValue = 0;
            S.firePercent = 0.0f; S.policePercent = 0.0f;
            S.roadPercent = yum > 0 ? ((float)yum) / ((float)S.roadFund) : 0.0f;
        }
    } else {
        S.fireValue = 0; S.policeValue = 0; S.roadValue = 0;
        S.firePercent = 1.0f; S.policePercent = 1.0f; S.roadPercent = 1.0f;
    }
    MP_NOUNROLL for (;;) {
        if (!S.autoBudget) {
            // showBudgetWindowAndStartWaiting is a front-end callback (no state)
            S.fireSpend = S.fireValue; S.policeSpend = S.policeValue; S.roadSpend = S.roadValue;
            total = S.fireSpend + S.policeSpend + S.roadSpend;
            int64_t moreDough = S.taxFund - total;
            S.totalFunds = (int)(S.totalFunds - (int)(-moreDough));    // spend(int) -> setFunds(int)
            return;
        }
        if (yum > total) {
            int64_t moreDough = S.taxFund - total;
            S.totalFunds = (int)(S.totalFunds - (int)(-moreDough));
            S.fireSpend = S.fireFund; S.policeSpend = S.policeFund; S.roadSpend = S.roadFund;
            return;
        }
        S.autoBudget = 0;
    }
}

// [LEAD] simulate.cpp:838-884 collectTax
MPN void collect_tax(Env &e)
{
    MP_COV(collect_tax);
    MP_TAB float RLevels[3] = { 0.7f, 0.9f, 1.2f }, FLevels[3] = { 1.4f, 1.2f, 0.8f };
    S.cashFlow = 0;
    if (!S.taxFlag) {
        S.cityTaxAverage = 0;
        S.policeFund = (int64_t)S.policeStationPop * 100;
        S.fireFund = (int64_t)S.fireStationPop * 100;
        S.roadFund = (int64_t)((S.roadTotal + (S.railTotal * 2)) * RLevels[S.gameLevel]);
        S.taxFund = (int64_t)((((int64_t)S.totalPop * S.landValueAverage) / 120) * S.cityTax * FLevels[S.gameLevel]);
        if (S.totalPop > 0) {
            S.cashFlow = (int16_t)(S.taxFund - (S.policeFund + S.fireFund + S.roadFund));
            do_budget(e);
        } else {
            S.roadEffect = MAX_ROAD_EFFECT;
            S.policeEffect = MAX_POLICE_EFFECT;
            S.fireEffect = MAX_FIRE_EFFECT;
        }
    }
}

// ------------------------------------------------------------------ phase 10
// [ALL] simulate.cpp:327-350 decTrafficMap (four cells per 32-bit word on the device)
MPD uint8_t dec_traffic_cell(int z)
{
    MP_COV(dec_traffic_cell);
    if (z == 0) return 0;
    if (z <= 24) return 0;
    return (uint8_t)(z > 200 ? z - 34 : z - 24);
}
MPN void dec_traffic_map(Env &e)
{
    MP_COV(dec_traffic_map);
#if defined(__CUDA_ARCH__)
    // 3000 cells (+8 zero pad bytes) = 188 uint4; four cells per 32-bit word with byte-SIMD:
    // z - (z > 200 ? 34 : 24) saturated at 0 is exactly the reference's three-way rule.
    uint4 *q = reinterpret_cast<uint4 *>(e.traffic());
    uint4 v[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int i = e.c.tid + k * 32;
        v[k] = (i < 188) ? q[i] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int i = e.c.tid + k * 32;
        if (i < 188 && (v[k].x | v[k].y | v[k].z | v[k].w)) {
            uint4 r;
            r.x = __vsubus4(v[k].x, (__vcmpgtu4(v[k].x, 0xc8c8c8c8u) & 0x0a0a0a0au) + 0x18181818u);
            r.y = __vsubus4(v[k].y, (__vcmpgtu4(v[k].y, 0xc8c8c8c8u) & 0x0a0a0a0au) + 0x18181818u);
            r.z = __vsubus4(v[k].z, (__vcmpgtu4(v[k].z, 0xc8c8c8c8u) & 0x0a0a0a0au) + 0x18181818u);
            r.w = __vsubus4(v[k].w, (__vcmpgtu4(v[k].w, 0xc8c8c8c8u) & 0x0a0a0a0au) + 0x18181818u);
            q[i] = r;
        }
    }
#else
    for (int i = e.c.tid; i < N2; i += e.c.n) e.traffic()[i] = dec_traffic_cell(e.traffic()[i]);
#endif
}

// [ALL] simulate.cpp:356-385 decRateOfGrowthMap
MPN void dec_rog_map(Env &e)
{
    MP_COV(dec_rog_map);
    for (int i = e.c.tid; i < N8; i += e.c.n) {
        int16_t z = e.rog()[i];
        if (z == 0) continue;
        if (z > 0) z--; else z++;
        e.rog()[i] = tclamp(z, (int16_t)-200, (int16_t)200);
    }
}

MPD int64_t get_population(const Env &e) { return ((int64_t)S.resPop + ((int64_t)S.comPop + S.indPop) * 8L) * 20L; }
MPD int city_class_of(int64_t p)                                    // evaluate.cpp:190-218
{
    MP_COV(city_class_of);
    int c = 0;
    if (p > 2000) c = 1;
    if (p > 10000) c = 2;
    if (p > 50000) c = 3;
    if (p > 100000) c = 4;
    if (p > 500000) c = 5;
    return c;
}

// [LEAD] message.cpp:77-231 sendMessages + :235-283 checkGrowth.  Messages themselves are
// front-end callbacks; only the state they gate (growth caps, category bookkeeping) is kept.
MPN void send_messages(Env &e)
{
    MP_COV(send_messages);
    if ((S.cityTime & 3) == 0) {
        int16_t category = 0;
        int64_t thisCityPop = get_population(e);
        if (S.cityPopLast > 0) {
            int lastClass = city_class_of(S.cityPopLast), newClass = city_class_of(thisCityPop);
            if (lastClass != newClass) {
                // MESSAGE_REACHED_TOWN.. = 35..39 (text.h); only equality with categoryLast matters
                if (newClass >= 1) category = (int16_t)(34 + newClass);
            }
        }
        if (category > 0 && category != S.categoryLast) S.categoryLast = category;
        S.cityPopLast = thisCityPop;
    }
    S.totalZonePop = (int16_t)(S.resZonePop + S.comZonePop + S.indZonePop);
    switch (S.cityTime & 63) {
    case 26: S.resCap = (S.resPop > 500 && S.stadiumPop == 0); break;
    case 28: S.indCap = (S.indPop > 70 && S.seaportPop == 0); break;
    case 30: S.comCap = (S.comPop > 100 && S.airportPop == 0); break;
    default: break;
    }
}

// ------------------------------------------------------------------ phase 11: power
// [LEAD] power.cpp:76-156 doPowerScan stack walk (powerGridMap already cleared by ALL).
// Literal replay: numPower counts visits AND pops, so the cut-off depends on traversal order.
// Returns true if the walk stayed well clear of the capacity cut-off (the result then does not depend
// on how many duplicate plant entries the stack held).
MPN bool power_scan_walk(Env &e)
{
    MP_COV(power_scan_walk);
    const int maxPower = S.coalPowerPop * 700 + S.nuclearPowerPop * 2000;     // <= 32767 * 2000 < 2^31
    int numPower = 0, sp = S.powerStackPointer;
    const bool no_seeds = sp <= 0;                // nothing to walk from: the grid stays empty
    // index form of Position::move for the four cardinal directions (N, E, S, W)
    while (sp > 0) {
        int p = e.pstack()[sp];
        sp--;
        int step = 0, conNum;
        do {
            numPower++;
            if (numPower > maxPower) { S.powerStackPointer = sp; return false; }
            p += step;
            e.power()[p >> 5] |= 1u << (p & 31);
            const int x = p / WORLD_H, y = p - x * WORLD_H;
            conNum = 0;
#define MP_PW_TRY(cond, off)                                                                          \
            if (conNum < 2 && (cond)) {                                                                   \
                const int q = p + (off);                                                                  \
                if ((e.map()[q] & CONDBIT) && !((e.power()[q >> 5] >> (q & 31)) & 1u)) { conNum++; step = (off); } \
            }
            MP_PW_TRY(y > 0, -1)
            MP_PW_TRY(x < WORLD_W - 1, WORLD_H)
            MP_PW_TRY(y < WORLD_H - 1, 1)
            MP_PW_TRY(x > 0, -WORLD_H)
#undef MP_PW_TRY
            if (conNum > 1 && sp < POWER_STACK_SIZE - 2) {
                sp++;
                e.pstack()[sp] = (uint16_t)p;
            }
        } while (conNum);
    }
    S.powerStackPointer = sp;
    return no_seeds || numPower + 256 < maxPower;
}

#if MP_WARP_COOP
// [ALL] doPowerScan as a lane-parallel flood fill over bitmaps, for the case in which the literal walk's
// only order-dependent feature -- the numPower cut-off -- cannot trigger.  numPower counts one step per
// newly powered tile plus one per pop, and a tile is pushed at most three times (once per neighbour it
// still has to leave through), so numPower <= 4 T + sp for T conductive tiles and sp seeds; with
// 4 T + sp + 256 < maxPower the walk powers exactly the conductive tiles 4-connected to a seed, which is
// what the fixpoint below computes.  Returns false (nothing done) when that bound does not hold or the
// grid is too small to be worth it; the caller then runs the literal walk.
//   C (conductive bitmap, scratch in tmp1) is built from the active bitmap; P = Env::power.
//   Tile i's neighbours are i -/+ 1 (same column: y -/+ 1, not across a column end) and i -/+ 100.
MPN bool power_scan_parallel(Env &e)
{
    MP_COV(power_scan_parallel);
    const int lane = e.c.tid;
    uint32_t *C = reinterpret_cast<uint32_t *>(e.tmp1());
    int T = 0, wlo = POWER_WORDS, whi = -1;
    MP_NOUNROLL for (int w = lane; w < POWER_WORDS; w += 32) {
        uint32_t a = e.act()[w], c = 0;
        while (a) {
            const int b = __ffs(a) - 1;
            a &= a - 1;
            if (e.map()[w * 32 + b] & CONDBIT) c |= 1u << b;
        }
        C[w] = c;
        if (c) { T += __popc(c); wlo = min(wlo, w); whi = max(whi, w); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        T += __shfl_xor_sync(0xffffffffu, T, o);
        wlo = min(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
        whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
    }
    const int sp = S.powerStackPointer;
    const int maxPower = S.coalPowerPop * 700 + S.nuclearPowerPop * 2000;
    if (T < 24 || 4 * T + sp + 256 >= maxPower) return false;
    e.c.sync();
    if (lane == 0)
        MP_NOUNROLL for (int k = 1; k <= sp; k++) {
            const int p = e.pstack()[k];
            e.power()[p >> 5] |= 1u << (p & 31);
            wlo = min(wlo, p >> 5); whi = max(whi, p >> 5);
        }
    wlo = __shfl_sync(0xffffffffu, wlo, 0);
    whi = __shfl_sync(0xffffffffu, whi, 0);
    e.c.sync();
    for (;;) {
        bool changed = false;
        MP_NOUNROLL for (int w = wlo + lane; w <= whi; w += 32) {
            const uint32_t c = C[w];
            if (!c) continue;
            uint32_t p = e.power()[w];
            // column ends inside this word: bit b with (w * 32 + b) % 100 == 0 has no upper neighbour, == 99 no lower
            const int r = (w * 32) % WORLD_H;
            const int b0 = (WORLD_H - r) % WORLD_H, b99 = WORLD_H - 1 - r;
            const uint32_t y0 = b0 < 32 ? 1u << b0 : 0u, y99 = b99 < 32 ? 1u << b99 : 0u;
            const uint32_t pm1 = w > 0 ? e.power()[w - 1] : 0u, pp1 = w < POWER_WORDS - 1 ? e.power()[w + 1] : 0u;
            const uint32_t pm3 = w >= 3 ? e.power()[w - 3] : 0u, pm4 = w >= 4 ? e.power()[w - 4] : 0u;
            const uint32_t pp3 = w + 3 < POWER_WORDS ? e.power()[w + 3] : 0u, pp4 = w + 4 < POWER_WORDS ? e.power()[w + 4] : 0u;
            uint32_t reach = ((((p << 1) | (pm1 >> 31)) & ~y0) | (((p >> 1) | (pp1 << 31)) & ~y99)
                              | (pm3 << 4) | (pm4 >> 28) | (pp3 >> 4) | (pp4 << 28));
            uint32_t add = reach & c & ~p;
            while (add) {                                   // run up and down the columns inside the word
                p |= add;
                add = (((p << 1) & ~y0) | ((p >> 1) & ~y99)) & c & ~p;
                changed = true;
            }
            e.power()[w] = p;
        }
        e.c.sync();
        if (!__any_sync(0xffffffffu, changed)) break;
    }
    if (lane == 0) S.powerStackPointer = 0;
    e.c.sync();
    return true;
}
#endif

// [ALL]
// set_new_power: every caller in the reference sets newPower = true right after doPowerScan
// (simulate.cpp:226, 289, 418); it is folded in here so that refresh_need sees the new value.
MPN void do_power_scan(Env &e, bool set_new_power = true)
{
    MP_COV(do_power_scan);
    // Memo (Scalars::powerQuiet / powerPure): nothing conductive has changed since a scan whose seeds were
    // all current plants, so the walk would rebuild the grid it already has; only its side effects remain --
    // the drained stack and newPower.
    e.c.sync();
    if (S.powerQuiet && S.powerPure && S.newPower) {
        e.c.sync();
        if (e.c.lead()) S.powerStackPointer = 0;
        e.c.sync();
        return;
    }
    for (int i = e.c.tid; i < POWER_WORDS; i += e.c.n) e.power()[i] = 0;
    e.c.sync();
    bool filled = false;
#if MP_WARP_COOP
    if (S.powerStackPointer > 0) filled = power_scan_parallel(e);
#endif
    if (e.c.lead()) {
        const bool headroom = filled ? true : power_scan_walk(e);
        S.powerPure = (S.powerQuiet && headroom) ? 1 : 0;
        S.powerQuiet = headroom ? 1 : 0;          // a cut-off walk can leave seeds on the stack for the next scan
        if (set_new_power) S.newPower = 1;
    }
    e.c.sync();
    refresh_need(e);        // inert building tiles must be revisited where PWRBIT no longer matches
}

// ------------------------------------------------------------------ smoothing passes
// [ALL] scan.cpp:551-575 smoothDitherMap, non-dither branch (donDither = 0 on this path): z = (sum of the four
// neighbours + self) >> 2, capped at 255 -- in sparse form.  The half-resolution maps of a gym city are zero outside the few dozen
// cells around the build window, and smoothing zeros gives zeros: `box` = {x0, x1, y0, y1} (inclusive)
// bounds every non-zero cell of src; the pass zero-fills dst, computes only the box grown by one cell,
// and grows `box` for the next pass.  An empty box (x0 > x1) means src is all zero.
struct CellBox { int x0, x1, y0, y1; };
MPD void zero_byte2(Env &e, uint8_t *m)
{
    MP_COV(zero_byte2);
    uint4 *q = reinterpret_cast<uint4 *>(m);                    // 3000 bytes + padding = 188 uint4
    for (int i = e.c.tid; i < (N2 + 15) / 16; i += e.c.n) q[i] = make_uint4(0, 0, 0, 0);
}
MPD void smooth_byte2_box(Env &e, const uint8_t *src, uint8_t *dst, CellBox &box)
{
    MP_COV(smooth_byte2_box);
    zero_byte2(e, dst);
    e.c.sync();
    if (box.x0 > box.x1) return;
    box.x0 = tmax(box.x0 - 1, 0); box.x1 = tmin(box.x1 + 1, W2 - 1);
    box.y0 = tmax(box.y0 - 1, 0); box.y1 = tmin(box.y1 + 1, H2 - 1);
    const int bh = box.y1 - box.y0 + 1, cells = (box.x1 - box.x0 + 1) * bh;
    for (int k = e.c.tid; k < cells; k += e.c.n) {
        const int cx = k / bh, x = box.x0 + cx, y = box.y0 + (k - cx * bh), i = x * H2 + y;
        int z = 0;
        if (x > 0) z += src[i - H2];
        if (x < W2 - 1) z += src[i + H2];
        if (y > 0) z += src[i - 1];
        if (y < H2 - 1) z += src[i + 1];
        z = (z + src[i]) >> 2;
        dst[i] = (uint8_t)tmin(z, 255);
    }
    e.c.sync();
}
// bounding box of the cells a lane has seen -> the group's box
MPD void box_add(CellBox &b, int x, int y)
{
    MP_COV(box_add);
    b.x0 = tmin(b.x0, x); b.x1 = tmax(b.x1, x); b.y0 = tmin(b.y0, y); b.y1 = tmax(b.y1, y);
}
MPD void box_reduce(Env &e, CellBox &b)
{
    MP_COV(box_reduce);
#if MP_WARP_COOP
    for (int o = 16; o > 0; o >>= 1) {
        b.x0 = min(b.x0, __shfl_xor_sync(0xffffffffu, b.x0, o));
        b.x1 = max(b.x1, __shfl_xor_sync(0xffffffffu, b.x1, o));
        b.y0 = min(b.y0, __shfl_xor_sync(0xffffffffu, b.y0, o));
        b.y1 = max(b.y1, __shfl_xor_sync(0xffffffffu, b.y1, o));
    }
#else
    (void)e; (void)b;
#endif
}

// [ALL] scan.cpp:80-105 smoothStationMap (in place through a temporary copy)
MPN void smooth_station(Env &e, int16_t *m)
{
    MP_COV(smooth_station);
    for (int i = e.c.tid; i < N8; i += e.c.n) e.stmp()[i] = m[i];
    e.c.sync();
    for (int i = e.c.tid; i < N8; i += e.c.n) {
        int x = i / H8, y = i - x * H8;
        int16_t edge = 0;
        if (x > 0) edge = (int16_t)(edge + e.stmp()[i - H8]);
        if (x < W8 - 1) edge = (int16_t)(edge + e.stmp()[i + H8]);
        if (y > 0) edge = (int16_t)(edge + e.stmp()[i - 1]);
        if (y < H8 - 1) edge = (int16_t)(edge + e.stmp()[i + 1]);
        edge = (int16_t)(e.stmp()[i] + edge / 4);
        m[i] = (int16_t)(edge / 2);
    }
    e.c.sync();
}

MPD int city_center_distance(const Env &e, int x, int y)            // scan.cpp:379-396
{
    MP_COV(city_center_distance);
    int xd = x > S.cityCenterX ? x - S.cityCenterX : S.cityCenterX - x;
    int yd = y > S.cityCenterY ? y - S.cityCenterY : S.cityCenterY - y;
    return tmin(xd + yd, 64);
}

MPD int pollution_value(int loc)                                    // scan.cpp:331-370
{
    MP_COV(pollution_value);
    if (loc < T_POWERBASE) {
        if (loc >= T_HTRFBASE) return 75;
        if (loc >= T_LTRFBASE) return 50;
        if (loc < T_ROADBASE) {
            if (loc > T_FIREBASE) return 90;
            if (loc >= T_RADTILE) return 255;
        }
        return 0;
    }
    if (loc <= T_LASTIND) return 0;
    if (loc < T_PORTBASE) return 50;
    if (loc <= T_LASTPOWERPLANT) return 100;
    return 0;
}

// block-wide integer sum of per-thread partials into e.c.red[slot] (valid after the barrier)
MPD void coop_add(Env &e, int slot, int v)
{
    MP_COV(coop_add);
#if MP_WARP_COOP
    MP_NOUNROLL for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (e.c.lead()) e.c.red[slot] += v;
#else
    e.c.red[slot] += v;
#endif
}

// [ALL] The reference's running-maximum sweeps (scan.cpp:283-303 pollution, :417-437 crime):
//   for each cell in order: if (z > max || (z == max && (getRandom16() & 3) == 0)) { max = z; site = cell; }
// Which cells draw depends only on the values (a cell draws iff it equals the maximum of
// everything before it), so 32 cells are handled at once: exclusive prefix maximum across the
// lanes, draws handed out in cell order by LCG jump-ahead, the last winning lane sets the site.
// `vals`: one byte per cell, `valid`: optional per-cell flag (crime counts only cells with land
// value), returns the winning cell index or -1; *count / *total are the census sums.
// [lo, hi): cell range that can hold live cells (everything outside is known to be dead); chunks keep their
// alignment to multiples of 32 so that the order of the draws is the full scan's.
MPD int tie_max_scan(Env &e, const uint8_t *vals, const uint8_t *valid, bool nonzero_only, int *count, int *total,
                     int lo = 0, int hi = N2)
{
    MP_COV(tie_max_scan);
#if MP_WARP_COOP
    const int lane = e.c.tid;
    const unsigned lt = (1u << lane) - 1u;
    int vmax = 0, site = -1, cnt = 0, tot = 0;
    uint32_t rng = S.rng;
    MP_NOUNROLL for (int base = lo & ~31; base < hi; base += 32) {
        const int i = base + lane;
        const bool in = i < N2;
        const int z = in ? vals[i] : 0;
        const bool live = in && (valid ? valid[i] != 0 : true) && (!nonzero_only || z != 0);
        const unsigned lm = __ballot_sync(0xffffffffu, live);
        if (!lm) continue;
        cnt += live ? 1 : 0;
        tot += live ? z : 0;
        // exclusive prefix maximum over live lanes, seeded with the running maximum
        int pm = live ? z : -1;
        MP_NOUNROLL for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, pm, o);
            if (lane >= o) pm = max(pm, y);
        }
        const int incl = pm;                                   // inclusive max of lanes <= me
        int excl = __shfl_up_sync(0xffffffffu, incl, 1);
        excl = lane == 0 ? -1 : excl;
        const int before = max(vmax, excl);
        const bool greater = live && z > before;
        const bool tie = live && z == before;
        const unsigned tm = __ballot_sync(0xffffffffu, tie);
        bool win = greater;
        if (tm) {
            if (tie) {
                uint32_t x = rng;
                const int k = __popc(tm & lt);
                MP_NOUNROLL for (int q = 0; q <= k; q++) x = x * 1103515245u + 12345u;
                win = ((((x & 0xffff00u) >> 8) & 0xffff) & 3) == 0;
            }
            const int nd = __popc(tm);
            MP_NOUNROLL for (int q = 0; q < nd; q++) rng = rng * 1103515245u + 12345u;
        }
        const unsigned wm = __ballot_sync(0xffffffffu, win);
        if (wm) site = base + (31 - __clz(wm));
        vmax = max(vmax, __shfl_sync(0xffffffffu, incl, 31));
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    if (lane == 0) S.rng = rng;
    __syncwarp();
    *count = cnt; *total = tot;
    return site;
#else
    int vmax = 0, site = -1, cnt = 0, tot = 0;
    for (int i = lo; i < tmin(hi, N2); i++) {
        const int z = vals[i];
        if ((valid && !valid[i]) || (nonzero_only && !z)) continue;
        cnt++; tot += z;
        if (z > vmax || (z == vmax && (rnd16(e) & 3) == 0)) { vmax = z; site = i; }
    }
    *count = cnt; *total = tot;
    return site;
#endif
}

// [ALL] scan.cpp:217-322 pollutionTerrainLandValueScan (+ :454-497 smoothTerrain, non-dither)
MPN void ptlv_scan(Env &e)
{
    MP_COV(ptlv_scan);
    if (e.c.lead()) { e.c.red[0] = 0; e.c.red[1] = 0; }
    // tempMap3: +15 per natural tile (0 < tile < RUBBLE) in each 4x4 block (Byte wrap)
    for (int i = e.c.tid; i < N4; i += e.c.n) {
        int bx = (i / H4) * 4, by = (i % H4) * 4, cnt = 0;
#if defined(__CUDA_ARCH__)
        // the four tiles of a block column are 8 aligned bytes (by is a multiple of 4, a map column is 200 bytes):
        // one 64-bit load per column, 1 <= id <= RUBBLE - 1 tested per halfword (this loop was the hottest line of
        // the kernel outside the barrier: 12 000 scalar tile loads per scan, profiles/r02_notes.md)
#pragma unroll
        for (int dx = 0; dx < 4; dx++) {
            const uint2 v = *reinterpret_cast<const uint2 *>(e.map() + tix(bx + dx, by));
            const unsigned a = v.x & 0x03ff03ffu, b = v.y & 0x03ff03ffu;
            cnt += (int)(((a & 0xffffu) - 1u) < (unsigned)(T_RUBBLE - 1)) + (int)(((a >> 16) - 1u) < (unsigned)(T_RUBBLE - 1))
                 + (int)(((b & 0xffffu) - 1u) < (unsigned)(T_RUBBLE - 1)) + (int)(((b >> 16) - 1u) < (unsigned)(T_RUBBLE - 1));
        }
#else
        MP_NOUNROLL for (int dx = 0; dx < 4; dx++)
            MP_NOUNROLL for (int dy = 0; dy < 4; dy++) {
                int loc = e.map()[tix(bx + dx, by + dy)] & LOMASK;
                if (loc && loc < T_RUBBLE) cnt++;
            }
#endif
        e.tmp3()[i] = (uint8_t)(cnt * 15);
    }
    e.c.sync();
    int lvtot = 0, lvnum = 0;
    CellBox pbox = { W2, -1, H2, -1 };               // cells with pollution sources
    for (int i = e.c.tid; i < N2; i += e.c.n) {
        int x = i / H2, y = i - x * H2, wx = x * 2, wy = y * 2;
        int plevel = 0;
        bool lvflag = false;
        // only tiles at or above FLOOD pollute or carry land value (getPollutionValue is 0 below
        // RADTILE, the flag needs ROADBASE), i.e. only tiles of the active bitmap; wy is even, so
        // the two tiles of a column are one aligned bit pair
        const int t0 = tix(wx, wy), t1 = t0 + WORLD_H;
        const unsigned pair = ((e.act()[t0 >> 5] >> (t0 & 31)) & 3u) | (((e.act()[t1 >> 5] >> (t1 & 31)) & 3u) << 2);
        if (pair) {
            MP_NOUNROLL for (int mx = wx; mx <= wx + 1; mx++)
                MP_NOUNROLL for (int my = wy; my <= wy + 1; my++) {
                    int loc = e.map()[tix(mx, my)] & LOMASK;
                    if (loc) {
                        if (loc < T_RUBBLE) continue;
                        plevel += pollution_value(loc);
                        if (loc >= T_ROADBASE) lvflag = true;
                    }
                }
        }
        plevel = tmin(plevel, 255);
        e.tmp1()[i] = (uint8_t)plevel;
        if (plevel) box_add(pbox, x, y);
        if (lvflag) {
            int dis = 34 - city_center_distance(e, wx, wy) / 2;
            dis = dis << 2;
            dis += e.terrain()[(x >> 1) * H4 + (y >> 1)];
            dis -= e.pollution()[i];
            if (e.crime()[i] > 190) dis -= 20;
            dis = tclamp(dis, 1, 250);
            e.landval()[i] = (uint8_t)dis;
            lvtot += dis;
            lvnum++;
        } else {
            e.landval()[i] = 0;
        }
    }
    coop_add(e, 0, lvtot);
    coop_add(e, 1, lvnum);
    e.c.sync();
    if (e.c.lead()) S.landValueAverage = e.c.red[1] > 0 ? (int16_t)(e.c.red[0] / e.c.red[1]) : (int16_t)0;
    box_reduce(e, pbox);
    smooth_byte2_box(e, e.tmp1(), e.tmp2(), pbox);
    smooth_byte2_box(e, e.tmp2(), e.tmp1(), pbox);
    // pollution max: draws depend on the running maximum, so this is the serial lane
    {
        const uint4 *q = reinterpret_cast<const uint4 *>(e.tmp1());
        uint4 *o = reinterpret_cast<uint4 *>(e.pollution());
        for (int i = e.c.tid; i < N2 / 16; i += e.c.n) o[i] = q[i];
        for (int i = (N2 / 16) * 16 + e.c.tid; i < N2; i += e.c.n) e.pollution()[i] = e.tmp1()[i];
    }
    e.c.sync();
    {
        int pnum, ptot;
        const int p_lo = pbox.x0 > pbox.x1 ? 0 : pbox.x0 * H2, p_hi = pbox.x0 > pbox.x1 ? 0 : (pbox.x1 + 1) * H2;
        const int site = tie_max_scan(e, e.tmp1(), nullptr, true, &pnum, &ptot, p_lo, p_hi);
        if (e.c.lead()) {
            if (site >= 0) {
                S.pollutionMaxX = (int16_t)((site / H2) * 2);
                S.pollutionMaxY = (int16_t)((site % H2) * 2);
            }
            S.pollutionAverage = pnum ? (int16_t)(ptot / pnum) : (int16_t)0;
        }
    }
    // smoothTerrain: tempMap3 -> terrainDensityMap
    for (int i = e.c.tid; i < N4; i += e.c.n) {
        int x = i / H4, y = i - x * H4;
        unsigned z = 0;
        if (x > 0) z += e.tmp3()[i - H4];
        if (x < W4 - 1) z += e.tmp3()[i + H4];
        if (y > 0) z += e.tmp3()[i - 1];
        if (y < H4 - 1) z += e.tmp3()[i + 1];
        e.terrain()[i] = (uint8_t)((uint8_t)(z / 4 + e.tmp3()[i]) / 2);
    }
    e.c.sync();
}

// [ALL] scan.cpp:403-450 crimeScan
MPN void crime_scan(Env &e)
{
    MP_COV(crime_scan);
    smooth_station(e, e.policeSt());
    smooth_station(e, e.policeSt());
    smooth_station(e, e.policeSt());
    // Only cells with land value have crime (everything else is zeroed), and land value exists only where
    // something is built: find the columns of the half-resolution map that hold any (128-bit loads, 16 cells
    // each), zero the two output maps, and visit only those columns.
    int c0 = W2, c1 = -1;
    {
        const uint4 *q = reinterpret_cast<const uint4 *>(e.landval());
        for (int j = e.c.tid; j < (N2 + 15) / 16; j += e.c.n) {
            const uint4 v = q[j];
            uint32_t any = v.x | v.y | v.z | v.w;
            if (j == (N2 + 15) / 16 - 1) any = v.x | v.y;        // the last word pair is padding (3000 = 187 * 16 + 8)
            if (any) { c0 = tmin(c0, (j * 16) / H2); c1 = tmax(c1, tmin((j * 16 + 15) / H2, W2 - 1)); }
        }
#if MP_WARP_COOP
        for (int o = 16; o > 0; o >>= 1) {
            c0 = min(c0, __shfl_xor_sync(0xffffffffu, c0, o));
            c1 = max(c1, __shfl_xor_sync(0xffffffffu, c1, o));
        }
#endif
    }
    zero_byte2(e, e.crime());
    zero_byte2(e, e.tmp2());
    e.c.sync();
    const int i_lo = c0 * H2, i_hi = (c1 + 1) * H2;                 // empty: i_lo > i_hi
    for (int i = i_lo + e.c.tid; i < i_hi; i += e.c.n) {
        int x = i / H2, y = i - x * H2;
        int z = e.landval()[i];
        if (z > 0) {
            z = 128 - z;
            z += e.pop()[i];
            z = tmin(z, 300);
            z -= e.policeSt()[(x >> 2) * H8 + (y >> 2)];
            z = tclamp(z, 0, 250);
            e.crime()[i] = (uint8_t)z;
            e.tmp2()[i] = 1;
        }
    }
    for (int i = e.c.tid; i < N8; i += e.c.n) e.policeEff()[i] = e.policeSt()[i];
    e.c.sync();
    {
        int numz, totz;
        const int site = tie_max_scan(e, e.crime(), e.tmp2(), false, &numz, &totz, tmin(i_lo, i_hi), i_hi);
        if (e.c.lead()) {
            if (site >= 0) {
                S.crimeMaxX = (int16_t)((site / H2) * 2);
                S.crimeMaxY = (int16_t)((site % H2) * 2);
            }
            S.crimeAverage = numz > 0 ? (int16_t)(totz / numz) : (int16_t)0;
        }
    }
    e.c.sync();
}

MPD int population_density_of(const Env &e, int x, int y, int tile)     // scan.cpp:186-210
{
    MP_COV(population_density_of);
    if (tile == T_FREEZ) return do_free_pop(e, x, y);
    if (tile < T_COMBASE) return res_zone_pop(tile);
    if (tile < T_INDBASE) return com_zone_pop(tile) * 8;
    if (tile < T_PORTBASE) return ind_zone_pop(tile) * 8;
    return 0;
}

// [ALL] scan.cpp:126-179 populationDensityScan (+ :580-590 computeComRateMap)
MPN void pop_density_scan(Env &e)
{
    MP_COV(pop_density_scan);
    if (e.c.lead()) { e.c.red[0] = 0; e.c.red[1] = 0; e.c.red[2] = 0; }
    e.c.sync();
    int xt = 0, yt = 0, zt = 0;
    CellBox box = { W2, -1, H2, -1 };
    // zone centres are active tiles: walk the active bitmap instead of all 12 000 tiles
    zero_byte2(e, e.tmp1());
    e.c.sync();
    for (int w = e.c.tid; w < POWER_WORDS; w += e.c.n) {
        uint32_t m = e.act()[w];
        while (m) {
            int b = 0;
#if defined(__CUDA_ARCH__)
            b = __ffs(m) - 1;
#else
            while (!((m >> b) & 1u)) b++;
#endif
            m &= m - 1;
            const int idx = w * 32 + b, mv = e.map()[idx];
            if (!(mv & ZONEBIT)) continue;
            const int x = idx / WORLD_H, y = idx - x * WORLD_H;
            xt += x; yt += y; zt++;
            // "last zone in scan order wins" inside a 2x2 cluster (tempMap1.worldSet overwrites)
            bool later = false;
            const int bx = x & ~1, by = y & ~1;
            MP_NOUNROLL for (int dx = 0; dx < 2; dx++)
                MP_NOUNROLL for (int dy = 0; dy < 2; dy++) {
                    const int o = tix(bx + dx, by + dy);
                    if (o > idx && (e.map()[o] & ZONEBIT)) later = true;
                }
            if (!later) {
                int pop = population_density_of(e, x, y, mv & LOMASK) * 8;
                e.tmp1()[(x >> 1) * H2 + (y >> 1)] = (uint8_t)tmin(pop, 254);
                if (pop) box_add(box, x >> 1, y >> 1);
            }
        }
    }
    coop_add(e, 0, xt);
    coop_add(e, 1, yt);
    coop_add(e, 2, zt);
    e.c.sync();
    box_reduce(e, box);
    smooth_byte2_box(e, e.tmp1(), e.tmp2(), box);
    smooth_byte2_box(e, e.tmp2(), e.tmp1(), box);
    smooth_byte2_box(e, e.tmp1(), e.tmp2(), box);
    {   // populationDensityMap = tempMap2 * 2 (Byte wrap), 16 cells per 128-bit word
        const uint4 *q = reinterpret_cast<const uint4 *>(e.tmp2());
        uint4 *o = reinterpret_cast<uint4 *>(e.pop());
        for (int i = e.c.tid; i < N2 / 16; i += e.c.n) {
            uint4 v = q[i];
            v.x = (v.x & 0x7f7f7f7fu) << 1; v.y = (v.y & 0x7f7f7f7fu) << 1;
            v.z = (v.z & 0x7f7f7f7fu) << 1; v.w = (v.w & 0x7f7f7f7fu) << 1;
            o[i] = v;
        }
        for (int i = (N2 / 16) * 16 + e.c.tid; i < N2; i += e.c.n) e.pop()[i] = (uint8_t)(e.tmp2()[i] * 2);
    }
    if (e.c.lead()) {
        if (e.c.red[2] > 0) {
            S.cityCenterX = (int16_t)(e.c.red[0] / e.c.red[2]);
            S.cityCenterY = (int16_t)(e.c.red[1] / e.c.red[2]);
        } else {
            S.cityCenterX = WORLD_W / 2;
            S.cityCenterY = WORLD_H / 2;
        }
    }
    e.c.sync();
    // computeComRateMap uses the NEW city centre only from the next scan on: the reference calls
    // it before updating cityCenterX/Y (scan.cpp:166-174), so it must see the old centre.
}

// [ALL] helper so that computeComRateMap sees the centre the reference saw
MPN void com_rate_map(Env &e, int cx, int cy)
{
    MP_COV(com_rate_map);
    for (int i = e.c.tid; i < N8; i += e.c.n) {
        int x = (i / H8) * 8, y = (i % H8) * 8;
        int xd = x > cx ? x - cx : cx - x, yd = y > cy ? y - cy : cy - y;
        int16_t z = (int16_t)(tmin(xd + yd, 64) / 2);
        z = (int16_t)(z * 4);
        e.comRate()[i] = (int16_t)(64 - z);
    }
}

// [ALL] scan.cpp:110-120 fireAnalysis
MPN void fire_analysis(Env &e)
{
    MP_COV(fire_analysis);
    smooth_station(e, e.fireSt());
    smooth_station(e, e.fireSt());
    smooth_station(e, e.fireSt());
    for (int i = e.c.tid; i < N8; i += e.c.n) e.fireEff()[i] = e.fireSt()[i];
    e.c.sync();
}

// ------------------------------------------------------------------ evaluation (evaluate.cpp)
// [LEAD] evaluate.cpp:100-124.  problemTable has PROBNUM+1 live slots because voteProblems walks
// problem = 0..PROBNUM inclusive (:293); slot PROBNUM reads as 0 in the pinned oracle (see
// oracle/build_ref.sh) -- it consumes a draw per visit and never votes.
MPN void city_evaluation(Env &e)
{
    MP_COV(city_evaluation);
    if (S.totalPop > 0) {
        int16_t pt[PROBNUM + 1];
        MP_NOUNROLL for (int z = 0; z <= PROBNUM; z++) pt[z] = 0;
        // getAssessedValue :154-169
        int64_t z = S.roadTotal * 5;
        z += S.railTotal * 10; z += S.policeStationPop * 1000; z += S.fireStationPop * 1000;
        z += S.hospitalPop * 400; z += S.stadiumPop * 3000; z += S.seaportPop * 5000;
        z += S.airportPop * 10000; z += S.coalPowerPop * 3000; z += S.nuclearPowerPop * 6000;
        S.cityAssessedValue = z * 1000;
        // doPopNum :176-187
        int64_t old = S.cityPop;
        S.cityPop = get_population(e);
        if (old == -1) old = S.cityPop;
        S.cityPopDelta = S.cityPop - old;
        S.cityClass = city_class_of(S.cityPop);
        // doProblems :227-269
        pt[0] = S.crimeAverage;
        pt[1] = S.pollutionAverage;
        pt[2] = (int16_t)(S.landValueAverage * 7 / 10);
        pt[3] = (int16_t)(S.cityTax * 10);
        {   // getTrafficAverage :305-327
            int64_t tt = 0;
            int16_t count = 1;
            MP_NOUNROLL for (int i = 0; i < N2; i++)
                if (e.landval()[i] > 0) { tt += e.traffic()[i]; count++; }
            S.trafficAverage = d2s((double)(tt / count) * 2.4);
            pt[4] = S.trafficAverage;
        }
        {   // getUnemployment :335-350
            int16_t b = (int16_t)((S.comPop + S.indPop) * 8);
            if (b == 0) pt[5] = 0;
            else {
                float r = ((float)S.resPop) / b;
                b = f2s((r - 1) * 255);
                pt[5] = tmin(b, (int16_t)255);
            }
        }
        int16_t fireSeverity = (int16_t)tmin(S.firePop * 5, 255);
        pt[6] = fireSeverity;
        // voteProblems :278-299
        MP_NOUNROLL for (int i = 0; i < PROBNUM; i++) S.problemVotes[i] = 0;
        int problem = 0, voteCount = 0, loopCount = 0;
        while (voteCount < 100 && loopCount < 600) {
            if (rnd(e, 300) < pt[problem]) {
                S.problemVotes[problem]++;     // problem == PROBNUM never gets here (pt = 0)
                voteCount++;
            }
            problem++;
            if (problem > PROBNUM) problem = 0;
            loopCount++;
        }
        bool taken[PROBNUM];
        MP_NOUNROLL for (int i = 0; i < PROBNUM; i++) taken[i] = false;
        MP_NOUNROLL for (int c = 0; c < PROBLEM_COMPLAINTS; c++) {
            int maxVotes = 0, best = NUMPROBLEMS;
            MP_NOUNROLL for (int i = 0; i < NUMPROBLEMS; i++)
                if (S.problemVotes[i] > maxVotes && !taken[i]) { best = i; maxVotes = S.problemVotes[i]; }
            S.problemOrder[c] = (int16_t)best;
            if (best < NUMPROBLEMS) taken[best] = true;
        }
        // getScore :359-451 (int, float and double mixed exactly as written there)
        int16_t last = S.cityScore;
        int x = 0;
        MP_NOUNROLL for (int i = 0; i < NUMPROBLEMS; i++) x += pt[i];
        x = x / 3;
        x = tmin(x, 256);
        int sc = tclamp((256 - x) * 4, 0, 1000);
        if (S.resCap) sc = (int)(sc * .85);
        if (S.comCap) sc = (int)(sc * .85);
        if (S.indCap) sc = (int)(sc * .85);
        if (S.roadEffect < MAX_ROAD_EFFECT) sc -= (int)(MAX_ROAD_EFFECT - S.roadEffect);
        if (S.policeEffect < MAX_POLICE_EFFECT)
            sc = (int)(sc * (0.9 + (S.policeEffect / (10.0001 * MAX_POLICE_EFFECT))));
        if (S.fireEffect < MAX_FIRE_EFFECT)
            sc = (int)(sc * (0.9 + (S.fireEffect / (10.0001 * MAX_FIRE_EFFECT))));
        if (S.resValve < -1000) sc = (int)(sc * .85);
        if (S.comValve < -1000) sc = (int)(sc * .85);
        if (S.indValve < -1000) sc = (int)(sc * .85);
        float SM = 1.0f;
        if (S.cityPop == 0 || S.cityPopDelta == 0) SM = 1.0f;
        else if (S.cityPopDelta == S.cityPop) SM = 1.0f;
        else if (S.cityPopDelta > 0) SM = ((float)S.cityPopDelta / S.cityPop) + 1.0f;
        else if (S.cityPopDelta < 0) SM = 0.95f + ((float)S.cityPopDelta / (S.cityPop - S.cityPopDelta));
        sc = (int)(sc * SM);
        sc = sc - fireSeverity - S.cityTax;
        float TM = (float)(S.unpoweredZoneCount + S.poweredZoneCount);
        if (TM > 0.0f) sc = (int)(sc * (float)(S.poweredZoneCount / TM));
        sc = tclamp(sc, 0, 1000);
        S.cityScore = (int16_t)((S.cityScore + sc) / 2);
        S.cityScoreDelta = (int16_t)(S.cityScore - last);
        // doVotes :454-464
        S.cityYes = 0;
        MP_NOUNROLL for (int i = 0; i < 100; i++)
            if (rnd(e, 1000) < S.cityScore) S.cityYes++;
    } else {
        // evalInit :129-147
        S.cityPop = 0; S.cityPopDelta = 0; S.cityAssessedValue = 0; S.cityClass = 0;
        S.cityScore = 500; S.cityScoreDelta = 0;
        MP_NOUNROLL for (int i = 0; i < PROBNUM; i++) S.problemVotes[i] = 0;
        MP_NOUNROLL for (int i = 0; i < PROBLEM_COMPLAINTS; i++) S.problemOrder[i] = NUMPROBLEMS;
        S.cityYes = 50;
    }
}

// ------------------------------------------------------------------ the 16-phase cycle
// value of lane 0 for every lane
MPD int coop_bcast(Env &e, int v)
{
    MP_COV(coop_bcast);
#if MP_WARP_COOP
    return __shfl_sync(0xffffffffu, v, 0);
#else
    (void)e;
    return v;
#endif
}
// scan periods by speed index (simulate.cpp:124-135); si == 2 (simSpeed 3) is the gym path
MPD bool scan_due(int cyc, int si, int slow, int mid, int fast)
{
    MP_COV(scan_due);
    // cyc % period == 0 without a runtime remainder (the compiler selects the period first and then emits its generic
    // 32-bit remainder -- int-to-float, reciprocal, two corrections -- or calls the 16-bit subroutine, five times a
    // cycle): simCycle < 1024 and every period <= 20, so floor(c / d) = (c * ceil(2^16 / d)) >> 16 exactly
    // (c * (d * ceil(2^16 / d) - 2^16) < 1024 * 20 < 2^16), with the reciprocals folded from the literals
    const unsigned c = (unsigned)cyc;
    const unsigned d = si == 2 ? (unsigned)fast : (si == 1 ? (unsigned)mid : (unsigned)slow);
    const unsigned r = si == 2 ? (65535u + (unsigned)fast) / (unsigned)fast
                               : (si == 1 ? (65535u + (unsigned)mid) / (unsigned)mid : (65535u + (unsigned)slow) / (unsigned)slow);
    return c - ((c * r) >> 16) * d == 0u;
}

// [ALL] one phase of simulate() (simulate.cpp:122-268).  `phase` is uniform across the lanes.
MPD void simulate_phase(Env &e, int phase)
{
    MP_COV(simulate_phase);
    // (speed index and cycle count are read by the phases that use them: phases 1..9 pay nothing for them)
#define MP_SI tclamp((int)S.simSpeed - 1, 0, 2)
    switch (phase) {
    case 0:
        if (e.c.lead()) {
            if (++S.simCycle > 1023) S.simCycle = 0;
            if (S.doInitialEval) {
                S.doInitialEval = 0;
                city_evaluation(e);
            }
            S.cityTime++;
            S.cityTaxAverage = (int16_t)(S.cityTaxAverage + S.cityTax);
            if (!(S.simCycle & 1)) set_valves(e);
            clear_census_scalars(e);
        }
#if MP_WARP_COOP
        {   // 2 x 195 int16 (padded to 400 bytes each in EnvLayout): one 128-bit store per lane and map
            static_assert(a16(N8 * 2) == 25 * 16, "station maps are cleared with 25 uint4 stores");
            const int i = e.c.tid;
            if (i < 25) {
                reinterpret_cast<uint4 *>(e.fireSt())[i] = make_uint4(0u, 0u, 0u, 0u);
                reinterpret_cast<uint4 *>(e.policeSt())[i] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
#else
        for (int i = e.c.tid; i < N8; i += e.c.n) { e.fireSt()[i] = 0; e.policeSt()[i] = 0; }
#endif
        break;
    case 1: case 2: case 3: case 4: case 5: case 6: case 7: case 8:
        map_scan(e, (phase - 1) * WORLD_W / 8, phase * WORLD_W / 8);
        break;
    case 9: {
        const int ct = (int)((uint32_t)S.cityTime % 48u);  // cityTime >= 0 and far below 2^32 on any rollout
        if ((ct & 3) == 0) { take10_census(e); e.c.sync(); }       // 48 = 4 * 12
        if (ct == 0) { take120_census(e); e.c.sync(); }
        if (e.c.lead() && ct == 0) {
            collect_tax(e);
            city_evaluation(e);
        }
        break;
    }
    case 10:
        if (!(S.simCycle % 5)) dec_rog_map(e);
        dec_traffic_map(e);
        if (e.c.lead()) send_messages(e);
        break;
    case 11:
        if (scan_due(S.simCycle, MP_SI, 2, 4, 5)) do_power_scan(e);          // sets newPower (see do_power_scan)
        break;
    case 12:
        if (scan_due(S.simCycle, MP_SI, 2, 7, 17)) ptlv_scan(e);
        break;
    case 13:
        if (scan_due(S.simCycle, MP_SI, 1, 8, 18)) crime_scan(e);
        break;
    case 14:
        if (scan_due(S.simCycle, MP_SI, 1, 9, 19)) {
            int cx = S.cityCenterX, cy = S.cityCenterY;
            e.c.sync();
            pop_density_scan(e);
            com_rate_map(e, cx, cy);
        }
        break;
    case 15:
        if (scan_due(S.simCycle, MP_SI, 1, 10, 20)) fire_analysis(e);
        {   // an earthquake (300-1000 random tile probes) is handed to all lanes instead of the serial lane
            int quake = 0;
            if (e.c.lead()) quake = do_disasters(e, true);
            if (coop_bcast(e, quake)) make_earthquake_coop(e);
        }
        break;
    }
#undef MP_SI
}

// [ALL] simulate.cpp:122-268 simulate
MPP void simulate(Env &e)
{
    MP_COV(simulate);
    const int phase = S.initSimLoad ? 0 : (S.phaseCycle & 15);
    e.c.sync();
    simulate_phase(e, phase);
    e.c.sync();
    if (e.c.lead()) S.phaseCycle = (int16_t)((phase + 1) & 15);
    e.c.sync();
}

// [ALL] simulate.cpp:98-118 simFrame + main.cpp:403-425 simLoop(true) (heatSteps = 0).
// One barrier per frame: the lead lane owns speedCycle / phaseCycle / spriteCycle, every lane
// reads the phase before the lead moves it on.
// The frame comes in two halves so that a CTA of several envs can put a barrier between the
// simulation phase and the sprite pass (each half then runs the same code in every warp).
// Returns phase + 1 if the phase was simulated, 0 if the speed divider skipped it.
// [ALL] one simulation phase + its closing barrier, out of line once (the frame loops call it)
MPR void sim_phase(Env &e, int phase)
{
    MP_COV(sim_phase);
    simulate_phase(e, phase);
    e.c.sync();
}
MPN int sim_loop_head(Env &e)
{
    MP_COV(sim_loop_head);
    const int phase = S.initSimLoad ? 0 : (S.phaseCycle & 15);
    int run = 0;
    if (e.c.lead()) {
        const int speed = S.simSpeed;
        if (speed != 0) {
            int sc = S.speedCycle + 1;
            if (sc > 1023) sc = 0;
            S.speedCycle = (int16_t)sc;
            run = speed >= 3 || !((speed == 1 && (sc % 5) != 0) || (speed == 2 && (sc % 3) != 0));
        }
    }
    run = coop_bcast(e, run);
    if (run) sim_phase(e, phase);
    return run ? phase + 1 : 0;
}
MPN void sim_loop_tail(Env &e, int ran)
{
    MP_COV(sim_loop_tail);
    if (e.c.lead()) {
        if (ran) S.phaseCycle = (int16_t)(ran & 15);
        if (S.spriteHead < 0) { if (S.simSpeed) S.spriteCycle++; }      // moveObjects with an empty list
        else move_objects(e);
    }
    e.c.sync();
}
MPD void sim_loop(Env &e)
{
    MP_COV(sim_loop);
    const int ran = sim_loop_head(e);
    sim_loop_tail(e, ran);
}

// [ALL] simulate.cpp:276-305 doSimInit (+ :390-421 initSimMemory)
MPD void do_sim_init(Env &e)
{
    MP_COV(do_sim_init);
    if (e.c.lead()) {
        S.phaseCycle = 0;
        S.simCycle = 0;
    }
    e.c.sync();
    if (S.initSimLoad == 2) {
        if (e.c.lead()) {
            // setCommonInits -> evalInit
            S.cityYes = 0; S.cityPop = 0; S.cityPopDelta = 0; S.cityAssessedValue = 0;
            S.cityClass = 0; S.cityScore = 500; S.cityScoreDelta = 0;
            MP_NOUNROLL for (int i = 0; i < PROBNUM; i++) S.problemVotes[i] = 0;
            MP_NOUNROLL for (int i = 0; i < PROBLEM_COMPLAINTS; i++) S.problemOrder[i] = NUMPROBLEMS;
            S.roadEffect = MAX_ROAD_EFFECT; S.policeEffect = MAX_POLICE_EFFECT;
            S.fireEffect = MAX_FIRE_EFFECT; S.taxFlag = 0; S.taxFund = 0;
            S.crimeRamp = 0; S.pollutionRamp = 0; S.totalPop = 0;
            S.resValve = 0; S.comValve = 0; S.indValve = 0;
            S.resCap = 0; S.comCap = 0; S.indCap = 0;
            S.externalMarket = 6.0f;
            S.disasterEvent = 0; S.scoreType = 0;
            S.powerStackPointer = 0;
        }
        for (int i = e.c.tid; i < 240; i += e.c.n) {
            e.resHist()[i] = 0; e.comHist()[i] = 0; e.indHist()[i] = 0;
            e.moneyHist()[i] = 128; e.crimeHist()[i] = 0; e.pollutionHist()[i] = 0;
        }
        e.c.sync();
        do_power_scan(e);
        if (e.c.lead()) { S.newPower = 1; S.initSimLoad = 0; }
        e.c.sync();
    }
    // initSimLoad == 1 (city loaded from file) is not reachable on the gym path
    if (e.c.lead()) { set_valves(e); clear_census_scalars(e); }
    for (int i = e.c.tid; i < N8; i += e.c.n) { e.fireSt()[i] = 0; e.policeSt()[i] = 0; }
    e.c.sync();
    map_scan(e, 0, WORLD_W);
    do_power_scan(e);
    if (e.c.lead()) S.newPower = 1;
    e.c.sync();
    ptlv_scan(e);
    crime_scan(e);
    {
        int cx = S.cityCenterX, cy = S.cityCenterY;
        e.c.sync();
        pop_density_scan(e);
        com_rate_map(e, cx, cy);
        e.c.sync();
    }
    fire_analysis(e);
    if (e.c.lead()) { S.totalPop = 1; S.doInitialEval = 1; }
    e.c.sync();
}

// [ALL] animate.cpp:204-219 animateTiles.  Every tile whose animation successor differs from itself
// has an id >= FIRE, so animated tiles are a subset of the active bitmap: each lane walks the set
// bits of the words it owns (and refreshes `need` for the tiles it changes; the id stays >= FIRE so
// `act` does not change).
MPN void animate_tiles(Env &e)
{
    MP_COV(animate_tiles);
    MP_NOUNROLL for (int w = e.c.tid; w < POWER_WORDS; w += e.c.n) {
        uint32_t m = e.act()[w], nd = e.need()[w];
        while (m) {
            int b = 0;
#if defined(__CUDA_ARCH__)
            b = __ffs(m) - 1;
#else
            while (!((m >> b) & 1u)) b++;
#endif
            m &= m - 1;
            const int i = w * 32 + b;
            int v = e.map()[i];
            if (!(v & ANIMBIT)) continue;
            const int nv = e.anim[v & LOMASK] | (v & ALLBITS);
            if (nv == v) continue;
            e.map()[i] = (uint16_t)nv;
            nd = tile_needs_visit(e, i, nv) ? (nd | (1u << b)) : (nd & ~(1u << b));
        }
        e.need()[w] = nd;
    }
    e.c.sync();
}

#undef S
} // namespace mp
