// mp_tools.h -- the editing tools (toolDown) and the terrain generator (reset path).
// [LEAD] throughout: tools draw from the env's RNG stream (putRubble, putDownPark) and terrain
// generation is one long RNG-ordered walk.
//
// ToolEffects (tool.h:90-170, tool.cpp:78-171) is a pending-modification overlay that reads
// through to the map and is applied only if funds suffice.  Every tile a single tool call can
// touch lies within [x-4, x+6] x [y-4, y+6] of the tool position, so the overlay here is a dense
// 16x16 window anchored at (x-5, y-5) with a presence mask (Env::ovl: 256 values + 8 mask words).
#pragma once
#include "mp_sim.h"

namespace mp {
#define S (*e.s)

enum ToolResult : int { TR_NO_MONEY = -2, TR_NEED_BULLDOZE = -1, TR_FAILED = 0, TR_OK = 1 };
enum EditingTool : int {
    TOOL_RESIDENTIAL = 0, TOOL_COMMERCIAL, TOOL_INDUSTRIAL, TOOL_FIRESTATION, TOOL_POLICESTATION,
    TOOL_QUERY, TOOL_WIRE, TOOL_BULLDOZER, TOOL_RAILROAD, TOOL_ROAD, TOOL_STADIUM, TOOL_PARK,
    TOOL_SEAPORT, TOOL_COALPOWER, TOOL_NUCLEARPOWER, TOOL_AIRPORT, TOOL_NETWORK, TOOL_WATER,
    TOOL_LAND, TOOL_FOREST, TOOL_COUNT
};

struct Fx {                 // one tool call's pending effects
    Env &e;
    int ox, oy, cost;
    bool any;               // something has been written: until then reads go straight to the map and the
                            // presence mask is not even cleared (most clear-before-build calls fail early)
    MPD Fx(Env &env, int x, int y) : e(env), ox(x - 5), oy(y - 5), cost(0), any(false) {}
    MPD int slot(int x, int y) const
    {
        int dx = x - ox, dy = y - oy;
        return (dx >= 0 && dx < 16 && dy >= 0 && dy < 16) ? dx * 16 + dy : -1;
    }
    MPD int value(int x, int y) const
    {
        if (!any) return tget(e, x, y);
        int s = slot(x, y);
        const uint32_t *m = (const uint32_t *)(e.ovl + 256);
        if (s >= 0 && ((m[s >> 5] >> (s & 31)) & 1u)) return e.ovl[s];
        return tget(e, x, y);
    }
    MPD int tile(int x, int y) const { return value(x, y) & LOMASK; }
    MPD void set(int x, int y, int v)
    {
        int s = slot(x, y);
        if (s < 0 || !inb(x, y)) return;
        uint32_t *m = (uint32_t *)(e.ovl + 256);
        if (!any) {
            MP_NOUNROLL for (int i = 0; i < 8; i++) m[i] = 0;
            any = true;
        }
        m[s >> 5] |= 1u << (s & 31);
        e.ovl[s] = (uint16_t)v;
    }
    MPD void apply()            // ToolEffects::modifyWorld (tool.cpp:113-137)
    {
        e.s->totalFunds = (int)(e.s->totalFunds - cost);
        if (!any) return;
        const uint32_t *m = (const uint32_t *)(e.ovl + 256);
        MP_NOUNROLL for (int w = 0; w < 8; w++) {
            uint32_t bits = m[w];
            while (bits) {                  // only the slots that were written
                int b = 0;
#if defined(__CUDA_ARCH__)
                b = __ffs(bits) - 1;
#else
                while (!((bits >> b) & 1u)) b++;
#endif
                bits &= bits - 1;
                const int s = w * 32 + b;
                mset(e, tix(ox + (s >> 4), oy + (s & 15)), e.ovl[s]);
            }
        }
    }
};

MPD int neutralize_road(int t) { return (t >= 64 && t <= 207) ? (t & 0xf) + 64 : t; }   // connect.cpp:107-113

MPD bool tally(int t)                                              // tool.cpp:381-388
{
    MP_COV(tally);
    return (t >= T_FIRSTRIVEDGE && t <= T_LASTRUBBLE) || (t >= T_POWERBASE + 2 && t <= T_POWERBASE + 12)
        || (t >= T_TINYEXP && t <= T_LASTTINYEXP + 2);
}
MPD bool is_tree(int v) { int t = v & LOMASK; return t >= T_WOODS_LOW && t <= T_WOODS_HIGH; }

// connect.cpp:554-734 fixSingle
MPN void fix_single(Fx &fx, int x, int y)
{
    MP_COV(fix_single);
    MP_TAB uint16_t RoadTable[16] = { T_ROADS, T_ROADS2, T_ROADS, T_ROADS3, T_ROADS2, T_ROADS2, T_ROADS4,
        T_ROADS8, T_ROADS, T_ROADS6, T_ROADS, T_ROADS7, T_ROADS5, T_ROADS10, T_ROADS9, T_INTERSECTION };
    MP_TAB uint16_t RailTable[16] = { T_LHRAIL, T_LVRAIL, T_LHRAIL, T_LVRAIL2, T_LVRAIL, T_LVRAIL,
        T_LVRAIL3, T_LVRAIL7, T_LHRAIL, T_LVRAIL5, T_LHRAIL, T_LVRAIL6, T_LVRAIL4, T_LVRAIL9, T_LVRAIL8,
        T_LVRAIL10 };
    MP_TAB uint16_t WireTable[16] = { T_LHPOWER, T_LVPOWER, T_LHPOWER, T_LVPOWER2, T_LVPOWER, T_LVPOWER,
        T_LVPOWER3, T_LVPOWER7, T_LHPOWER, T_LVPOWER5, T_LHPOWER, T_LVPOWER6, T_LVPOWER4, T_LVPOWER9,
        T_LVPOWER8, T_LVPOWER10 };
    int adj = 0;
    int tile = neutralize_road(fx.tile(x, y));
    if (tile >= T_ROADS && tile <= T_INTERSECTION) {
        MP_NOUNROLL for (int k = 0; k < 4; k++) {
            bool vert = (k == 0 || k == 2);
            int nx = x + (k == 1) - (k == 3), ny = y - (k == 0) + (k == 2);
            bool in = (k == 0) ? y > 0 : (k == 1) ? x < WORLD_W - 1 : (k == 2) ? y < WORLD_H - 1 : x > 0;
            if (!in) continue;
            int t = neutralize_road(fx.tile(nx, ny));
            if (vert) {
                if ((t == T_HRAILROAD || (t >= T_ROADBASE && t <= T_VROADPOWER)) && t != T_HROADPOWER
                    && t != T_VRAILROAD && t != T_ROADBASE)
                    adj |= 1 << k;
            } else {
                if ((t == T_VRAILROAD || (t >= T_ROADBASE && t <= T_VROADPOWER)) && t != T_VROADPOWER
                    && t != T_HRAILROAD && t != T_VBRIDGE)
                    adj |= 1 << k;
            }
        }
        fx.set(x, y, RoadTable[adj] | BULLBIT | BURNBIT);
        return;
    }
    if (tile >= T_LHRAIL && tile <= T_LVRAIL10) {
        MP_NOUNROLL for (int k = 0; k < 4; k++) {
            bool vert = (k == 0 || k == 2);
            int nx = x + (k == 1) - (k == 3), ny = y - (k == 0) + (k == 2);
            bool in = (k == 0) ? y > 0 : (k == 1) ? x < WORLD_W - 1 : (k == 2) ? y < WORLD_H - 1 : x > 0;
            if (!in) continue;
            int t = neutralize_road(fx.tile(nx, ny));
            if (t >= T_RAILHPOWERV && t <= T_VRAILROAD) {
                if (vert) { if (t != T_RAILHPOWERV && t != T_HRAILROAD && t != T_HRAIL) adj |= 1 << k; }
                else { if (t != T_RAILVPOWERH && t != T_VRAILROAD && t != T_VRAIL) adj |= 1 << k; }
            }
        }
        fx.set(x, y, RailTable[adj] | BULLBIT | BURNBIT);
        return;
    }
    if (tile >= T_LHPOWER && tile <= T_LVPOWER10) {
        MP_NOUNROLL for (int k = 0; k < 4; k++) {
            bool vert = (k == 0 || k == 2);
            int nx = x + (k == 1) - (k == 3), ny = y - (k == 0) + (k == 2);
            bool in = (k == 0) ? y > 0 : (k == 1) ? x < WORLD_W - 1 : (k == 2) ? y < WORLD_H - 1 : x > 0;
            if (!in) continue;
            int v = fx.value(nx, ny);
            if (v & CONDBIT) {
                int t = neutralize_road(v & LOMASK);
                if (vert) { if (t != T_VPOWER && t != T_VROADPOWER && t != T_RAILVPOWERH) adj |= 1 << k; }
                else { if (t != T_HPOWER && t != T_HROADPOWER && t != T_RAILHPOWERV) adj |= 1 << k; }
            }
        }
        fx.set(x, y, WireTable[adj] | BLBNCNBIT);
    }
}

// fixSingle only ever redraws a road, rail or wire tile; the test is made at the call site so that the
// usual case (five tiles of ground, water or buildings) costs no calls
MPD bool fix_candidate(const Fx &fx, int x, int y)
{
    MP_COV(fix_candidate);
    const int t = neutralize_road(fx.tile(x, y));
    return (t >= T_ROADS && t <= T_INTERSECTION) || (t >= T_LHRAIL && t <= T_LVRAIL10)
        || (t >= T_LHPOWER && t <= T_LVPOWER10);
}
MPD void fix_zone(Fx &fx, int x, int y)                            // connect.cpp:523-545
{
    MP_COV(fix_zone);
    if (fix_candidate(fx, x, y)) fix_single(fx, x, y);
    if (y > 0 && fix_candidate(fx, x, y - 1)) fix_single(fx, x, y - 1);
    if (x < WORLD_W - 1 && fix_candidate(fx, x + 1, y)) fix_single(fx, x + 1, y);
    if (y < WORLD_H - 1 && fix_candidate(fx, x, y + 1)) fix_single(fx, x, y + 1);
    if (x > 0 && fix_candidate(fx, x - 1, y)) fix_single(fx, x - 1, y);
}

MPD int lay_doze(Fx &fx, int x, int y)                             // connect.cpp:193-236
{
    MP_COV(lay_doze);
    int v = fx.value(x, y);
    if (!(v & BULLBIT)) return TR_FAILED;
    int t = neutralize_road(v & LOMASK);
    switch (t) {
    case T_HBRIDGE: case T_VBRIDGE: case T_BRWV: case T_BRWH: case T_HBRDG0: case T_HBRDG1: case T_HBRDG2:
    case T_HBRDG3: case T_VBRDG0: case T_VBRDG1: case T_VBRDG2: case T_VBRDG3: case T_HPOWER: case T_VPOWER:
    case T_HRAIL: case T_VRAIL:
        fx.set(x, y, T_RIVER);
        break;
    default:
        fx.set(x, y, T_DIRT);
        break;
    }
    fx.cost += 1;
    return TR_OK;
}

MPN int lay_road(Fx &fx, int x, int y)                             // connect.cpp:246-329
{
    MP_COV(lay_road);
    int cost = 10, t = fx.tile(x, y);
    switch (t) {
    case T_DIRT: fx.set(x, y, T_ROADS | BULLBIT | BURNBIT); break;
    case T_RIVER: case T_REDGE: case T_CHANNEL: {
        cost = 50;
        bool done = false;
        if (x < WORLD_W - 1) {
            int n = neutralize_road(fx.tile(x + 1, y));
            if (n == T_VRAILROAD || n == T_HBRIDGE || (n >= T_ROADS && n <= T_HROADPOWER)) {
                fx.set(x, y, T_HBRIDGE | BULLBIT); done = true;
            }
        }
        if (!done && x > 0) {
            int n = neutralize_road(fx.tile(x - 1, y));
            if (n == T_VRAILROAD || n == T_HBRIDGE || (n >= T_ROADS && n <= T_INTERSECTION)) {
                fx.set(x, y, T_HBRIDGE | BULLBIT); done = true;
            }
        }
        if (!done && y < WORLD_H - 1) {
            int n = neutralize_road(fx.tile(x, y + 1));
            if (n == T_HRAILROAD || n == T_VROADPOWER || (n >= T_VBRIDGE && n <= T_INTERSECTION)) {
                fx.set(x, y, T_VBRIDGE | BULLBIT); done = true;
            }
        }
        if (!done && y > 0) {
            int n = neutralize_road(fx.tile(x, y - 1));
            if (n == T_HRAILROAD || n == T_VROADPOWER || (n >= T_VBRIDGE && n <= T_INTERSECTION)) {
                fx.set(x, y, T_VBRIDGE | BULLBIT); done = true;
            }
        }
        if (!done) return TR_FAILED;
        break;
    }
    case T_LHPOWER: fx.set(x, y, T_VROADPOWER | CONDBIT | BURNBIT | BULLBIT); break;
    case T_LVPOWER: fx.set(x, y, T_HROADPOWER | CONDBIT | BURNBIT | BULLBIT); break;
    case T_LHRAIL: fx.set(x, y, T_HRAILROAD | BURNBIT | BULLBIT); break;
    case T_LVRAIL: fx.set(x, y, T_VRAILROAD | BURNBIT | BULLBIT); break;
    default: return TR_FAILED;
    }
    fx.cost += cost;
    return TR_OK;
}

MPN int lay_rail(Fx &fx, int x, int y)                             // connect.cpp:338-425
{
    MP_COV(lay_rail);
    int cost = 20, t = neutralize_road(fx.tile(x, y));
    switch (t) {
    case T_DIRT: fx.set(x, y, T_LHRAIL | BULLBIT | BURNBIT); break;
    case T_RIVER: case T_REDGE: case T_CHANNEL: {
        cost = 100;
        bool done = false;
        if (x < WORLD_W - 1) {
            int n = neutralize_road(fx.tile(x + 1, y));
            if (n == T_RAILHPOWERV || n == T_HRAIL || (n >= T_LHRAIL && n <= T_HRAILROAD)) {
                fx.set(x, y, T_HRAIL | BULLBIT); done = true;
            }
        }
        if (!done && x > 0) {
            int n = neutralize_road(fx.tile(x - 1, y));
            if (n == T_RAILHPOWERV || n == T_HRAIL || (n > T_VRAIL && n < T_VRAILROAD)) {
                fx.set(x, y, T_HRAIL | BULLBIT); done = true;
            }
        }
        if (!done && y < WORLD_H - 1) {
            int n = neutralize_road(fx.tile(x, y + 1));
            if (n == T_RAILVPOWERH || n == T_VRAILROAD || (n > T_HRAIL && n < T_HRAILROAD)) {
                fx.set(x, y, T_VRAIL | BULLBIT); done = true;
            }
        }
        if (!done && y > 0) {
            int n = neutralize_road(fx.tile(x, y - 1));
            if (n == T_RAILVPOWERH || n == T_VRAILROAD || (n > T_HRAIL && n < T_HRAILROAD)) {
                fx.set(x, y, T_VRAIL | BULLBIT); done = true;
            }
        }
        if (!done) return TR_FAILED;
        break;
    }
    case T_LHPOWER: fx.set(x, y, T_RAILVPOWERH | CONDBIT | BURNBIT | BULLBIT); break;
    case T_LVPOWER: fx.set(x, y, T_RAILHPOWERV | CONDBIT | BURNBIT | BULLBIT); break;
    case T_ROADS: fx.set(x, y, T_VRAILROAD | BURNBIT | BULLBIT); break;
    case T_ROADS2: fx.set(x, y, T_HRAILROAD | BURNBIT | BULLBIT); break;
    default: return TR_FAILED;
    }
    fx.cost += cost;
    return TR_OK;
}

MPN int lay_wire(Fx &fx, int x, int y)                             // connect.cpp:434-520
{
    MP_COV(lay_wire);
    int cost = 5, t = neutralize_road(fx.tile(x, y));
    switch (t) {
    case T_DIRT: fx.set(x, y, T_LHPOWER | CONDBIT | BURNBIT | BULLBIT); break;
    case T_RIVER: case T_REDGE: case T_CHANNEL: {
        cost = 25;
        bool done = false;
        MP_NOUNROLL for (int k = 0; k < 4 && !done; k++) {
            bool horiz = k < 2;
            int nx = x + (k == 0) - (k == 1), ny = y + (k == 2) - (k == 3);
            bool in = (k == 0) ? x < WORLD_W - 1 : (k == 1) ? x > 0 : (k == 2) ? y < WORLD_H - 1 : y > 0;
            if (!in) continue;
            int v = fx.value(nx, ny);
            if (v & CONDBIT) {
                int n = neutralize_road(v & LOMASK);
                if (horiz) {
                    if (n != T_HROADPOWER && n != T_RAILHPOWERV && n != T_HPOWER) {
                        fx.set(x, y, T_VPOWER | CONDBIT | BULLBIT); done = true;
                    }
                } else {
                    if (n != T_VROADPOWER && n != T_RAILVPOWERH && n != T_VPOWER) {
                        fx.set(x, y, T_HPOWER | CONDBIT | BULLBIT); done = true;
                    }
                }
            }
        }
        if (!done) return TR_FAILED;
        break;
    }
    case T_ROADS: fx.set(x, y, T_HROADPOWER | CONDBIT | BURNBIT | BULLBIT); break;
    case T_ROADS2: fx.set(x, y, T_VROADPOWER | CONDBIT | BURNBIT | BULLBIT); break;
    case T_LHRAIL: fx.set(x, y, T_RAILHPOWERV | CONDBIT | BURNBIT | BULLBIT); break;
    case T_LVRAIL: fx.set(x, y, T_RAILVPOWERH | CONDBIT | BURNBIT | BULLBIT); break;
    default: return TR_FAILED;
    }
    fx.cost += cost;
    return TR_OK;
}

enum ConnectCmd : int { CC_FIX = 0, CC_BULLDOZE, CC_ROAD, CC_RAILROAD, CC_WIRE };

// connect.cpp:122-184 connectTile
MPN int connect_tile(Fx &fx, int x, int y, int cmd)
{
    MP_COV(connect_tile);
    Env &e = fx.e;
    int result = TR_OK;
    if (!inb(x, y)) return TR_FAILED;
    if ((cmd == CC_ROAD || cmd == CC_RAILROAD || cmd == CC_WIRE) && S.autoBulldoze) {
        int v = fx.value(x, y);
        if (v & BULLBIT) {
            v = neutralize_road(v & LOMASK);
            if ((v >= T_TINYEXP && v <= T_LASTTINYEXP) || (v < T_HBRIDGE && v != T_DIRT)) {
                fx.cost += 1;
                fx.set(x, y, T_DIRT);
            }
        }
    }
    switch (cmd) {
    case CC_FIX: break;
    case CC_BULLDOZE: result = lay_doze(fx, x, y); break;
    case CC_ROAD: result = lay_road(fx, x, y); break;
    case CC_RAILROAD: result = lay_rail(fx, x, y); break;
    case CC_WIRE: result = lay_wire(fx, x, y); break;
    }
    fix_zone(fx, x, y);
    return result;
}

// tool.cpp:299-378 checkBigZone: offset of the zone centre from a non-centre tile of a 4x4 / 6x6
MPD int check_big_zone(int id, int &dh, int &dv)
{
    MP_COV(check_big_zone);
    dh = 0; dv = 0;
    MP_TAB int base4[4] = { T_POWERPLANT, T_PORT, T_NUCLEAR, T_STADIUM };
    MP_NOUNROLL for (int k = 0; k < 4; k++) {
        int d = id - base4[k];
        if (d == 0) return 4;
        if (d == 1) { dh = -1; return 4; }
        if (d == 4) { dv = -1; return 4; }
        if (d == 5) { dh = -1; dv = -1; return 4; }
    }
    if (id >= T_COALSMOKE3 && id <= T_COALSMOKE3 + 2) { dh = -1; return 4; }
    int d = id - T_AIRPORT;
    if (d >= 0 && d <= 21 && (d % 6) <= 3) { dh = -(d % 6); dv = -(d / 6); return 6; }
    return 0;
}

MPD int check_size(int t)                                          // tool.cpp:397-413
{
    MP_COV(check_size);
    if ((t >= T_RESBASE - 1 && t <= T_PORTBASE - 1) || (t >= T_LASTPOWERPLANT + 1 && t <= T_POLICESTATION + 4)
        || (t >= T_CHURCH1BASE && t <= T_CHURCH7LAST))
        return 3;
    if ((t >= T_PORTBASE && t <= T_LASTPORT) || (t >= T_COALBASE && t <= T_LASTPOWERPLANT)
        || (t >= T_STADIUMBASE && t <= T_LASTZONE))
        return 4;
    return 0;
}

MPD void put_rubble(Fx &fx, int x, int y, int size)                // tool.cpp:910-924
{
    MP_COV(put_rubble);
    Env &e = fx.e;
    MP_NOUNROLL for (int xx = x; xx < x + size; xx++)
        MP_NOUNROLL for (int yy = y; yy < y + size; yy++)
            if (inb(xx, yy)) {
                int t = fx.tile(xx, yy);
                if (t != T_RADTILE && t != T_DIRT) {
                    t = S.doAnimation ? (T_TINYEXP + rnd(e, 2)) : (int)T_SOMETINYEXP;
                    fx.set(xx, yy, t | ANIMBIT | BULLBIT);
                }
            }
}

// tool.cpp:965-1072 bulldozerTool(x, y, effects)
MPN int bulldozer_tool(Fx &fx, int x, int y)
{
    MP_COV(bulldozer_tool);
    if (!inb(x, y)) return TR_FAILED;
    int v = fx.value(x, y), t = v & LOMASK, size, dh = 0, dv = 0;
    if (v & ZONEBIT) size = check_size(t);
    else size = check_big_zone(t, dh, dv);
    if (size > 0) {
        fx.cost += 1;
        put_rubble(fx, x + dh - 1, y + dv - 1, size);
        return TR_OK;
    }
    int result = connect_tile(fx, x, y, CC_BULLDOZE);
    if ((t == T_RIVER || t == T_REDGE || t == T_CHANNEL) && t != fx.tile(x, y)) fx.cost += 5;
    return result;
}

// tool.cpp:480-540 buildBuilding (+ :415-478 prepareBuildingSite / putBuilding / checkBorder)
MPN int build_building(Fx &fx, int x, int y, int size, int base, int toolCost, bool anim)
{
    MP_COV(build_building);
    Env &e = fx.e;
    x--; y--;
    if (x < 0 || x + size > WORLD_W) return TR_FAILED;
    if (y < 0 || y + size > WORLD_H) return TR_FAILED;
    MP_NOUNROLL for (int dy = 0; dy < size; dy++)
        MP_NOUNROLL for (int dx = 0; dx < size; dx++) {
            int t = fx.tile(x + dx, y + dy);
            if (t == T_DIRT) continue;
            if (!S.autoBulldoze) return TR_NEED_BULLDOZE;
            if (!tally(t)) return TR_NEED_BULLDOZE;
            fx.set(x + dx, y + dy, T_DIRT);
            fx.cost += 1;
        }
    fx.cost += toolCost;
    MP_NOUNROLL for (int dy = 0; dy < size; dy++)
        MP_NOUNROLL for (int dx = 0; dx < size; dx++) {
            int v = base | BNCNBIT;
            if (dx == 1) {
                if (dy == 1) v |= ZONEBIT;
                else if (dy == 2 && anim) v |= ANIMBIT;
            }
            fx.set(x + dx, y + dy, v);
            base++;
        }
    MP_NOUNROLL for (int c = 0; c < size; c++) connect_tile(fx, x + c, y - 1, CC_FIX);
    MP_NOUNROLL for (int c = 0; c < size; c++) connect_tile(fx, x - 1, y + c, CC_FIX);
    MP_NOUNROLL for (int c = 0; c < size; c++) connect_tile(fx, x + c, y + size, CC_FIX);
    MP_NOUNROLL for (int c = 0; c < size; c++) connect_tile(fx, x + size, y + c, CC_FIX);
    return TR_OK;
}

// generate.cpp:436-475 smoothTreesAt(x, y, preserve, effects)
MPD void smooth_trees_at_fx(Fx &fx, int x, int y, bool preserve)
{
    MP_COV(smooth_trees_at_fx);
    MP_TAB int8_t dx[4] = { -1, 0, 1, 0 }, dy[4] = { 0, 1, 0, -1 };
    MP_TAB int8_t treeTable[16] = { 0, 0, 0, 34, 0, 0, 36, 35, 0, 32, 0, 33, 30, 31, 29, 37 };
    if (!is_tree(fx.value(x, y))) return;
    int bits = 0;
    MP_NOUNROLL for (int z = 0; z < 4; z++) {
        bits <<= 1;
        int xx = x + dx[z], yy = y + dy[z];
        if (inb(xx, yy) && is_tree(fx.value(xx, yy))) bits++;
    }
    int temp = treeTable[bits & 15];
    if (temp) {
        if (temp != T_WOODS && ((x + y) & 1)) temp -= 8;
        fx.set(x, y, temp | BLBNBIT);
    } else if (!preserve) {
        fx.set(x, y, temp);
    }
}

// [LEAD] tool.cpp:1375-1516 doTool + toolDown.  Returns the ToolResult.
MPN int tool_down(Env &e, int tool, int x, int y)
{
    MP_COV(tool_down);
    MP_TAB int16_t costOf[TOOL_COUNT] = { 100, 100, 100, 500, 500, 0, 5, 1, 20, 10, 5000, 10, 3000, 3000,
        5000, 10000, 100, 0, 0, 0 };
    x = (int16_t)x; y = (int16_t)y;
    Fx fx(e, x, y);
    int result;
    switch (tool) {
    case TOOL_RESIDENTIAL: result = build_building(fx, x, y, 3, T_RESBASE, costOf[tool], false); break;
    case TOOL_COMMERCIAL: result = build_building(fx, x, y, 3, T_COMBASE, costOf[tool], false); break;
    case TOOL_INDUSTRIAL: result = build_building(fx, x, y, 3, T_INDBASE, costOf[tool], false); break;
    case TOOL_FIRESTATION: result = build_building(fx, x, y, 3, T_FIRESTBASE, costOf[tool], false); break;
    case TOOL_POLICESTATION: result = build_building(fx, x, y, 3, T_POLICESTBASE, costOf[tool], false); break;
    case TOOL_STADIUM: result = build_building(fx, x, y, 4, T_STADIUMBASE, costOf[tool], false); break;
    case TOOL_SEAPORT: result = build_building(fx, x, y, 4, T_PORTBASE, costOf[tool], false); break;
    case TOOL_COALPOWER: result = build_building(fx, x, y, 4, T_COALBASE, costOf[tool], false); break;
    case TOOL_NUCLEARPOWER: result = build_building(fx, x, y, 4, T_NUCLEARBASE, costOf[tool], true); break;
    case TOOL_AIRPORT: result = build_building(fx, x, y, 6, T_AIRPORTBASE, costOf[tool], false); break;
    case TOOL_QUERY: return inb(x, y) ? TR_OK : TR_FAILED;
    case TOOL_WIRE: result = inb(x, y) ? connect_tile(fx, x, y, CC_WIRE) : (int)TR_FAILED; break;
    case TOOL_RAILROAD: result = inb(x, y) ? connect_tile(fx, x, y, CC_RAILROAD) : (int)TR_FAILED; break;
    case TOOL_ROAD: result = inb(x, y) ? connect_tile(fx, x, y, CC_ROAD) : (int)TR_FAILED; break;
    case TOOL_BULLDOZER: result = bulldozer_tool(fx, x, y); break;
    case TOOL_PARK:
        if (!inb(x, y)) { result = TR_FAILED; break; }
        {   // putDownPark tool.cpp:208-230 (the draw happens before the DIRT test)
            int16_t value = rnd(e, 4);
            int tile = BURNBIT | BULLBIT;
            if (value == 4) tile |= T_FOUNTAIN | ANIMBIT;
            else tile |= value + T_WOODS2;
            if (fx.value(x, y) != T_DIRT) { result = TR_NEED_BULLDOZE; break; }
            fx.cost += costOf[TOOL_PARK];
            fx.set(x, y, tile);
            result = TR_OK;
        }
        break;
    case TOOL_NETWORK:
        if (!inb(x, y)) { result = TR_FAILED; break; }
        {   // putDownNetwork tool.cpp:233-252
            int t = fx.tile(x, y);
            if (t != T_DIRT && tally(t)) {
                fx.cost += costOf[TOOL_BULLDOZER];
                fx.set(x, y, T_DIRT);
                t = T_DIRT;
            }
            if (t != T_DIRT) { result = TR_NEED_BULLDOZE; break; }
            fx.set(x, y, T_TELEBASE | CONDBIT | BURNBIT | BULLBIT | ANIMBIT);
            fx.cost += costOf[TOOL_NETWORK];
            result = TR_OK;
        }
        break;
    case TOOL_WATER:
        if (!inb(x, y)) { result = TR_FAILED; break; }
        result = bulldozer_tool(fx, x, y);
        if (result == TR_OK) {
            if (fx.tile(x, y) == T_RIVER) result = TR_FAILED;
            else fx.set(x, y, T_RIVER);
        }
        break;
    case TOOL_LAND:
        if (!inb(x, y)) { result = TR_FAILED; break; }
        bulldozer_tool(fx, x, y);
        if (fx.tile(x, y) == T_DIRT) result = TR_FAILED;
        else { fx.set(x, y, T_DIRT); result = TR_OK; }
        break;
    case TOOL_FOREST:
        if (!inb(x, y)) { result = TR_FAILED; break; }
        {
            result = TR_OK;
            int v = fx.value(x, y);
            if (is_tree(v)) break;
            if ((v & LOMASK) != T_DIRT) result = bulldozer_tool(fx, x, y);
            v = fx.value(x, y);
            if (v == T_DIRT) {
                fx.set(x, y, T_WOODS | BLBNBIT);
                MP_TAB int8_t dx[8] = { -1, 0, 1, -1, 1, -1, 0, 1 }, dy[8] = { -1, -1, -1, 0, 0, 1, 1, 1 };
                MP_NOUNROLL for (int i = 0; i < 8; i++)
                    if (inb(x + dx[i], y + dy[i])) smooth_trees_at_fx(fx, x + dx[i], y + dy[i], true);
                result = TR_OK;
            } else {
                result = TR_FAILED;
            }
        }
        break;
    default:
        return TR_FAILED;
    }
    if (result == TR_OK) {
        if (S.totalFunds < fx.cost) return TR_NO_MONEY;
        fx.apply();
    }
    return result;
}

// ------------------------------------------------------------------ terrain generator (generate.cpp)
MPD void put_on_map(Env &e, int ch, int x, int y)                  // generate.cpp:572-598
{
    MP_COV(put_on_map);
    if (ch == 0 || !inb(x, y)) return;
    int temp = e.map()[tix(x, y)];
    if (temp != T_DIRT) {
        temp &= LOMASK;
        if (temp == T_RIVER && ch != T_CHANNEL) return;
        if (temp == T_CHANNEL) return;
    }
    mset(e, tix(x, y), (uint16_t)ch);
}

// generate.cpp:604-626 plopBRiver: 9x9 disc, REDGE rim, RIVER body, CHANNEL centre
MPN void plop_b_river(Env &e, int px, int py)
{
    MP_COV(plop_b_river);
    MP_TAB int8_t lo[9] = { 3, 2, 1, 0, 0, 0, 1, 2, 3 };    // first non-empty column of each row
    MP_NOUNROLL for (int x = 0; x < 9; x++)
        MP_NOUNROLL for (int y = 0; y < 9; y++) {
            int l = lo[y], h = 8 - l, ch = 0;
            if (x >= l && x <= h) ch = (x == l || x == h || y == 0 || y == 8) ? T_REDGE : T_RIVER;
            if (x == 4 && y == 4) ch = T_CHANNEL;
            put_on_map(e, ch, px + x, py + y);
        }
}

// generate.cpp:632-651 plopSRiver: 6x6 disc
MPN void plop_s_river(Env &e, int px, int py)
{
    MP_COV(plop_s_river);
    MP_TAB int8_t lo[6] = { 2, 1, 0, 0, 1, 2 };
    MP_NOUNROLL for (int x = 0; x < 6; x++)
        MP_NOUNROLL for (int y = 0; y < 6; y++) {
            int l = lo[y], h = 5 - l, ch = 0;
            if (x >= l && x <= h) ch = (x == l || x == h || y == 0 || y == 5) ? T_REDGE : T_RIVER;
            put_on_map(e, ch, px + x, py + y);
        }
}

MPN int do_river(Env &e, int px, int py, int riverDir, int terrainDir, bool big)   // generate.cpp:498-570
{
    MP_COV(do_river);
    const int rate1 = 100, rate2 = 200;        // terrainCurveLevel = -1 (micropolis.cpp:393)
    const int ext = big ? 4 : 3;
    while (inb(px + ext, py + ext)) {
        if (big) plop_b_river(e, px, py); else plop_s_river(e, px, py);
        if (rnd(e, rate1) < 10) terrainDir = riverDir;
        else {
            if (rnd(e, rate2) > 90) terrainDir = rot45(terrainDir, 1);
            if (rnd(e, rate2) > 90) terrainDir = rot45(terrainDir, 7);
        }
        pos_move(px, py, terrainDir);
    }
    return terrainDir;
}

MPN void smooth_river(Env &e)                                      // generate.cpp:354-395
{
    MP_COV(smooth_river);
    MP_TAB int8_t dx[4] = { -1, 0, 1, 0 }, dy[4] = { 0, 1, 0, -1 };
    MP_TAB int16_t REdTab[16] = { 13 | BULLBIT, 13 | BULLBIT, 17 | BULLBIT, 15 | BULLBIT, 5 | BULLBIT, 2,
        19 | BULLBIT, 17 | BULLBIT, 9 | BULLBIT, 11 | BULLBIT, 2, 13 | BULLBIT, 7 | BULLBIT, 9 | BULLBIT,
        5 | BULLBIT, 2 };
    MP_NOUNROLL for (int x = 0; x < WORLD_W; x++)
        MP_NOUNROLL for (int y = 0; y < WORLD_H; y++)
            if (e.map()[tix(x, y)] == T_REDGE) {
                int bits = 0;
                MP_NOUNROLL for (int z = 0; z < 4; z++) {
                    bits <<= 1;
                    int xx = x + dx[z], yy = y + dy[z];
                    if (inb(xx, yy)) {
                        int t = e.map()[tix(xx, yy)] & LOMASK;
                        if (t != T_DIRT && (t < T_WOODS_LOW || t > T_WOODS_HIGH)) bits++;
                    }
                }
                int16_t temp = REdTab[bits & 15];
                if (temp != T_RIVER && rnd(e, 1)) temp++;
                mset(e, tix(x, y), (uint16_t)temp);
            }
}

MPN void smooth_trees(Env &e)                                      // generate.cpp:410-422 (+ :425-430)
{
    MP_COV(smooth_trees);
    MP_TAB int8_t dx[4] = { -1, 0, 1, 0 }, dy[4] = { 0, 1, 0, -1 };
    MP_TAB int8_t treeTable[16] = { 0, 0, 0, 34, 0, 0, 36, 35, 0, 32, 0, 33, 30, 31, 29, 37 };
    MP_NOUNROLL for (int x = 0; x < WORLD_W; x++)
        MP_NOUNROLL for (int y = 0; y < WORLD_H; y++)
            if (is_tree(e.map()[tix(x, y)])) {
                int bits = 0;
                MP_NOUNROLL for (int z = 0; z < 4; z++) {
                    bits <<= 1;
                    int xx = x + dx[z], yy = y + dy[z];
                    if (inb(xx, yy) && is_tree(e.map()[tix(xx, yy)])) bits++;
                }
                int temp = treeTable[bits & 15];
                if (temp) {
                    if (temp != T_WOODS && ((x + y) & 1)) temp -= 8;
                    mset(e, tix(x, y), (uint16_t)(temp | BLBNBIT));
                } else {
                    mset(e, tix(x, y), (uint16_t)temp);
                }
            }
}

MPN void do_trees(Env &e)                                          // generate.cpp:330-348 (+ :299-324)
{
    MP_COV(do_trees);
    int16_t amount = (int16_t)(rnd(e, 100) + 50);          // terrainTreeLevel = -1
    MP_NOUNROLL for (int16_t i = 0; i < amount; i++) {
        int16_t xloc = rnd(e, WORLD_W - 1), yloc = rnd(e, WORLD_H - 1);
        int16_t numTrees = (int16_t)(rnd(e, 150) + 50);
        int tx = xloc, ty = yloc;
        while (numTrees > 0) {
            int dir = D_N + rnd(e, 7);
            pos_move(tx, ty, dir);
            if (!inb(tx, ty)) break;
            if ((e.map()[tix(tx, ty)] & LOMASK) == T_DIRT) mset(e, tix(tx, ty), (uint16_t)(T_WOODS | BLBNBIT));
            numTrees--;
        }
    }
    smooth_trees(e);
    smooth_trees(e);
}

MPN void make_naked_island(Env &e)                                 // generate.cpp:193-231
{
    MP_COV(make_naked_island);
    MP_NOUNROLL for (int x = 0; x < WORLD_W; x++)
        MP_NOUNROLL for (int y = 0; y < WORLD_H; y++)
            mset(e, tix(x, y), (x < 5 || x >= WORLD_W - 5 || y < 5 || y >= WORLD_H - 5) ? T_RIVER : T_DIRT);
    MP_NOUNROLL for (int x = 0; x < WORLD_W - 5; x += 2) {
        int my = ernd(e, 18);
        plop_b_river(e, x, my);
        my = (WORLD_H - 10) - ernd(e, 18);
        plop_b_river(e, x, my);
        plop_s_river(e, x, 0);
        plop_s_river(e, x, WORLD_H - 6);
    }
    MP_NOUNROLL for (int y = 0; y < WORLD_H - 5; y += 2) {
        int mx = ernd(e, 18);
        plop_b_river(e, mx, y);
        mx = (WORLD_W - 10) - ernd(e, 18);
        plop_b_river(e, mx, y);
        plop_s_river(e, 0, y);
        plop_s_river(e, WORLD_W - 6, y);
    }
}

// [LEAD] generate.cpp:119-160 generateMap(seed) with the constructor's terrain settings
// (terrainTreeLevel = terrainLakeLevel = terrainCurveLevel = terrainCreateIsland = -1)
MPN void generate_map(Env &e, int seed)
{
    MP_COV(generate_map);
    S.rng = (uint32_t)seed;
    if (rnd(e, 100) < 10) {
        make_naked_island(e);
        smooth_river(e);
        do_trees(e);
        return;
    }
    MP_NOUNROLL for (int i = 0; i < NTILES; i++) mset(e, i, T_DIRT);
    {
        int xs = 40 + rnd(e, WORLD_W - 80), ys = 33 + rnd(e, WORLD_H - 67);
        // doRivers generate.cpp:478-495
        int riverDir = D_N + rnd(e, 3) * 2;
        do_river(e, xs, ys, riverDir, riverDir, true);
        riverDir = rot45(riverDir, 4);
        int terrainDir = do_river(e, xs, ys, riverDir, riverDir, true);
        riverDir = D_N + rnd(e, 3) * 2;
        do_river(e, xs, ys, riverDir, terrainDir, false);
    }
    {   // makeLakes generate.cpp:249-291
        int16_t numLakes = rnd(e, 10);
        while (numLakes > 0) {
            int x = rnd(e, WORLD_W - 21) + 10, y = rnd(e, WORLD_H - 20) + 10;
            int numPlops = rnd(e, 12) + 2;
            while (numPlops > 0) {
                // Position(pos, getRandom(12) - 6, getRandom(12) - 6): g++ evaluates the arguments
                // right to left, so the y offset is drawn first (pinned by the compiled oracle)
                int oy = rnd(e, 12) - 6;
                int ox = rnd(e, 12) - 6;
                if (rnd(e, 4)) plop_s_river(e, x + ox, y + oy);
                else plop_b_river(e, x + ox, y + oy);
                numPlops--;
            }
            numLakes--;
        }
    }
    smooth_river(e);
    do_trees(e);
}


#if MP_WARP_COOP
// ------------------------------------------------------------------ terrain generator, all 32 lanes
// The generator above is a serial program on the lead lane: ~1.5 M instructions, 36 ms per city on one lane
// (profiles/r01_launches.csv: k_mp_env_reset_begin = 250 ms for 32 768 cities, 13 env-steps' worth).  Its draws
// are few and its tile work is wide, so the same map is produced with the tile work spread over the warp and the
// draws handed out in reference order by LCG jump-ahead:
//   * a river / lake plop is 81 (36) distinct cells: one per lane, three (two) rounds;
//   * smoothRiver's result for an edge tile depends on whether its neighbours are "land" (dirt or woods), which
//     no rewrite of the pass changes -- every tile of a 32-tile word is computed at once, the coin flips
//     (getRandom(1), one draw each unless the 1-in-32768 rejection hits) ranked in scan order;
//   * a tree splash is a random walk whose path does not depend on the map: 32 steps at a time, positions by a
//     prefix sum of the step vectors, cut in front of the first step that touches the map border (Position::move
//     clamps and has its NORTH_EAST / SOUTH_WEST quirks there) or whose draw is rejected -- that one step is
//     replayed by the lead lane;
//   * smoothTrees reads the NEW value of its west and north neighbours and the OLD value of the other two: a
//     wavefront over the anti-diagonals x + y = d.
// Tiles are stored directly: generate_some_city rebuilds the derived bitmaps afterwards (rebuild_active).
MPD void gen_put(Env &e, int ch, int x, int y)                      // generate.cpp:572-598 putOnMap
{
    MP_COV(gen_put);
    if (ch == 0 || !inb(x, y)) return;
    int temp = e.map()[tix(x, y)];
    if (temp != T_DIRT) {
        temp &= LOMASK;
        if (temp == T_RIVER && ch != T_CHANNEL) return;
        if (temp == T_CHANNEL) return;
    }
    e.map()[tix(x, y)] = (uint16_t)ch;
}
MPD void gen_plop_b(Env &e, int px, int py)                         // generate.cpp:604-626
{
    MP_COV(gen_plop_b);
    MP_NOUNROLL for (int c = e.c.tid; c < 81; c += 32) {
        const int x = c / 9, y = c - x * 9;
        const int l = y < 4 ? 3 - y : (y > 4 ? y - 5 : 0);         // first non-empty column of the row
        const int h = 8 - l;
        int ch = 0;
        if (x >= l && x <= h) ch = (x == l || x == h || y == 0 || y == 8) ? T_REDGE : T_RIVER;
        if (x == 4 && y == 4) ch = T_CHANNEL;
        gen_put(e, ch, px + x, py + y);
    }
    __syncwarp();
}
MPD void gen_plop_s(Env &e, int px, int py)                         // generate.cpp:632-651
{
    MP_COV(gen_plop_s);
    MP_NOUNROLL for (int c = e.c.tid; c < 36; c += 32) {
        const int x = c / 6, y = c - x * 6;
        const int l = y < 2 ? 2 - y : (y > 3 ? y - 3 : 0);
        const int h = 5 - l;
        int ch = 0;
        if (x >= l && x <= h) ch = (x == l || x == h || y == 0 || y == 5) ? T_REDGE : T_RIVER;
        gen_put(e, ch, px + x, py + y);
    }
    __syncwarp();
}
MPN int gen_river(Env &e, int px, int py, int riverDir, int terrainDir, bool big)   // generate.cpp:498-570
{
    MP_COV(gen_river);
    const int ext = big ? 4 : 3;
    while (inb(px + ext, py + ext)) {
        if (big) gen_plop_b(e, px, py); else gen_plop_s(e, px, py);
        if (e.c.lead()) {
            if (rnd(e, 100) < 10) terrainDir = riverDir;
            else {
                if (rnd(e, 200) > 90) terrainDir = rot45(terrainDir, 1);
                if (rnd(e, 200) > 90) terrainDir = rot45(terrainDir, 7);
            }
            pos_move(px, py, terrainDir);
        }
        px = coop_bcast(e, px); py = coop_bcast(e, py); terrainDir = coop_bcast(e, terrainDir);
    }
    return terrainDir;
}

// generate.cpp:354-395 smoothRiver for one tile (lead lane: the tile whose coin flip was rejected)
MPD int gen_redge_value(Env &e, int idx)
{
    MP_COV(gen_redge_value);
    MP_TAB int16_t REdTab[16] = { 13 | BULLBIT, 13 | BULLBIT, 17 | BULLBIT, 15 | BULLBIT, 5 | BULLBIT, 2,
        19 | BULLBIT, 17 | BULLBIT, 9 | BULLBIT, 11 | BULLBIT, 2, 13 | BULLBIT, 7 | BULLBIT, 9 | BULLBIT,
        5 | BULLBIT, 2 };
    const int x = idx / WORLD_H, y = idx - x * WORLD_H;
    int bits = 0;
    MP_NOUNROLL for (int z = 0; z < 4; z++) {
        bits <<= 1;
        const int xx = x + (z == 2) - (z == 0), yy = y + (z == 1) - (z == 3);
        if (inb(xx, yy)) {
            const int t = e.map()[tix(xx, yy)] & LOMASK;
            if (t != T_DIRT && (t < T_WOODS_LOW || t > T_WOODS_HIGH)) bits++;
        }
    }
    return REdTab[bits & 15];
}
MPN void gen_smooth_river(Env &e)
{
    MP_COV(gen_smooth_river);
    const int lane = e.c.tid;
    MP_NOUNROLL for (int base = 0; base < NTILES; base += 32) {
        int pos = base;
        for (;;) {
            const int idx = base + lane;
            const bool mine = idx >= pos && e.map()[idx] == T_REDGE;
            if (!__any_sync(0xffffffffu, mine)) break;
            int temp = 0;
            if (mine) temp = gen_redge_value(e, idx);
            const bool draws = mine && temp != T_RIVER;
            const unsigned dm = __ballot_sync(0xffffffffu, draws);
            int r = 0;
            if (draws) r = lcg_r16(lcg_jump(S.rng, __popc(dm & ((1u << lane) - 1u)) + 1));
            const unsigned rej = __ballot_sync(0xffffffffu, draws && r >= 65534);      // getRandom(1): 65534 = 0xffff / 2 * 2
            const int stop = rej ? __ffs(rej) - 1 : 32;
            __syncwarp();                                                               // every neighbour read is done
            if (mine && lane < stop) e.map()[idx] = (uint16_t)(temp + (draws ? (r & 1) : 0));
            if (lane == 0) S.rng = lcg_jump(S.rng, __popc(stop < 32 ? (dm & ((1u << stop) - 1u)) : dm));
            __syncwarp();
            if (stop == 32) break;
            if (lane == 0) {                                                            // the rejected flip, literally
                int16_t t2 = (int16_t)gen_redge_value(e, base + stop);
                if (t2 != T_RIVER && rnd(e, 1)) t2++;
                e.map()[base + stop] = (uint16_t)t2;
            }
            __syncwarp();
            pos = base + stop + 1;
        }
    }
}

// generate.cpp:410-430 smoothTrees as a wavefront over the anti-diagonals
MPN void gen_smooth_trees(Env &e)
{
    MP_COV(gen_smooth_trees);
    MP_TAB int8_t treeTable[16] = { 0, 0, 0, 34, 0, 0, 36, 35, 0, 32, 0, 33, 30, 31, 29, 37 };
    const int lane = e.c.tid;
    MP_NOUNROLL for (int d = 0; d < WORLD_W + WORLD_H - 1; d++) {
        const int x0 = d < WORLD_H ? 0 : d - (WORLD_H - 1), x1 = d < WORLD_W ? d : WORLD_W - 1;
        MP_NOUNROLL for (int x = x0 + lane; x <= x1; x += 32) {
            const int y = d - x, idx = tix(x, y);
            if (!is_tree(e.map()[idx])) continue;
            int bits = 0;
            if (x > 0 && is_tree(e.map()[idx - WORLD_H])) bits |= 8;
            if (y < WORLD_H - 1 && is_tree(e.map()[idx + 1])) bits |= 4;
            if (x < WORLD_W - 1 && is_tree(e.map()[idx + WORLD_H])) bits |= 2;
            if (y > 0 && is_tree(e.map()[idx - 1])) bits |= 1;
            int temp = treeTable[bits];
            if (temp) {
                if (temp != T_WOODS && (d & 1)) temp -= 8;
                temp |= BLBNBIT;
            }
            e.map()[idx] = (uint16_t)temp;
        }
        __syncwarp();
    }
}

// generate.cpp:299-348 doTrees / treeSplash
MPN void gen_trees(Env &e)
{
    MP_COV(gen_trees);
    const int lane = e.c.tid;
    int amount = 0;
    if (lane == 0) amount = rnd(e, 100) + 50;
    amount = coop_bcast(e, amount);
    MP_NOUNROLL for (int i = 0; i < amount; i++) {
        int tx = 0, ty = 0, numTrees = 0;
        if (lane == 0) { tx = rnd(e, WORLD_W - 1); ty = rnd(e, WORLD_H - 1); numTrees = rnd(e, 150) + 50; }
        tx = coop_bcast(e, tx); ty = coop_bcast(e, ty); numTrees = coop_bcast(e, numTrees);
        while (numTrees > 0) {
            const int n = numTrees < 32 ? numTrees : 32;
            const int r = lcg_r16(lcg_jump(S.rng, lane + 1));
            const int dir = r & 7;                                     // getRandom(7) = r % 8 unless r >= 65528; D_N + dir
            // step vector of direction D_N + dir: N NE E SE S SW W NW
            const int dx = (dir >= 1 && dir <= 3) - (dir >= 5), dy = (dir >= 3 && dir <= 5) - (dir == 0 || dir == 1 || dir == 7);
            int sx = dx, sy = dy;
            MP_NOUNROLL for (int o = 1; o < 32; o <<= 1) {
                const int ax = __shfl_up_sync(0xffffffffu, sx, o), ay = __shfl_up_sync(0xffffffffu, sy, o);
                if (lane >= o) { sx += ax; sy += ay; }
            }
            const int x = tx + sx, y = ty + sy;
            const unsigned bad = __ballot_sync(0xffffffffu, lane < n && (r >= 65528 || !inb(x, y)));
            const int stop = bad ? __ffs(bad) - 1 : n;                 // steps [0, stop) are plain moves
            if (lane < stop && (e.map()[tix(x, y)] & LOMASK) == T_DIRT) e.map()[tix(x, y)] = (uint16_t)(T_WOODS | BLBNBIT);
            if (stop > 0) {
                tx = __shfl_sync(0xffffffffu, x, stop - 1);
                ty = __shfl_sync(0xffffffffu, y, stop - 1);
                if (lane == 0) S.rng = lcg_jump(S.rng, stop);
            }
            numTrees -= stop;
            __syncwarp();
            if (stop < n) {                                            // the step at the border / the rejected draw, literally
                if (lane == 0) {
                    const int d2 = D_N + rnd(e, 7);
                    pos_move(tx, ty, d2);
                    if ((e.map()[tix(tx, ty)] & LOMASK) == T_DIRT) e.map()[tix(tx, ty)] = (uint16_t)(T_WOODS | BLBNBIT);
                }
                tx = coop_bcast(e, tx); ty = coop_bcast(e, ty);
                numTrees--;
                __syncwarp();
            }
        }
    }
    gen_smooth_trees(e);
    gen_smooth_trees(e);
}

// [ALL] generate.cpp:119-160 generateMap(seed), same map and same final RNG state as generate_map above
MPN void generate_map_coop(Env &e, int seed)
{
    MP_COV(generate_map_coop);
    const int lane = e.c.tid;
    int island = 0;
    if (lane == 0) { S.rng = (uint32_t)seed; island = rnd(e, 100) < 10; }
    island = coop_bcast(e, island);
    if (island) {                                                      // makeNakedIsland generate.cpp:193-231
        MP_NOUNROLL for (int i = lane; i < NTILES; i += 32) {
            const int x = i / WORLD_H, y = i - x * WORLD_H;
            e.map()[i] = (x < 5 || x >= WORLD_W - 5 || y < 5 || y >= WORLD_H - 5) ? T_RIVER : T_DIRT;
        }
        __syncwarp();
        MP_NOUNROLL for (int x = 0; x < WORLD_W - 5; x += 2) {
            int a = 0, b = 0;
            if (lane == 0) { a = ernd(e, 18); b = (WORLD_H - 10) - ernd(e, 18); }
            a = coop_bcast(e, a); b = coop_bcast(e, b);
            gen_plop_b(e, x, a);
            gen_plop_b(e, x, b);
            gen_plop_s(e, x, 0);
            gen_plop_s(e, x, WORLD_H - 6);
        }
        MP_NOUNROLL for (int y = 0; y < WORLD_H - 5; y += 2) {
            int a = 0, b = 0;
            if (lane == 0) { a = ernd(e, 18); b = (WORLD_W - 10) - ernd(e, 18); }
            a = coop_bcast(e, a); b = coop_bcast(e, b);
            gen_plop_b(e, a, y);
            gen_plop_b(e, b, y);
            gen_plop_s(e, 0, y);
            gen_plop_s(e, WORLD_W - 6, y);
        }
    } else {
        MP_NOUNROLL for (int i = lane; i < NTILES; i += 32) e.map()[i] = T_DIRT;
        __syncwarp();
        int xs = 0, ys = 0, riverDir = 0;
        if (lane == 0) { xs = 40 + rnd(e, WORLD_W - 80); ys = 33 + rnd(e, WORLD_H - 67); riverDir = D_N + rnd(e, 3) * 2; }
        xs = coop_bcast(e, xs); ys = coop_bcast(e, ys); riverDir = coop_bcast(e, riverDir);
        gen_river(e, xs, ys, riverDir, riverDir, true);                // doRivers generate.cpp:478-495
        riverDir = rot45(riverDir, 4);
        const int terrainDir = gen_river(e, xs, ys, riverDir, riverDir, true);
        if (lane == 0) riverDir = D_N + rnd(e, 3) * 2;
        riverDir = coop_bcast(e, riverDir);
        gen_river(e, xs, ys, riverDir, terrainDir, false);
        int numLakes = 0;                                              // makeLakes generate.cpp:249-291
        if (lane == 0) numLakes = rnd(e, 10);
        numLakes = coop_bcast(e, numLakes);
        while (numLakes > 0) {
            int x = 0, y = 0, numPlops = 0;
            if (lane == 0) { x = rnd(e, WORLD_W - 21) + 10; y = rnd(e, WORLD_H - 20) + 10; numPlops = rnd(e, 12) + 2; }
            x = coop_bcast(e, x); y = coop_bcast(e, y); numPlops = coop_bcast(e, numPlops);
            while (numPlops > 0) {
                int oy = 0, ox = 0, small = 0;
                if (lane == 0) { oy = rnd(e, 12) - 6; ox = rnd(e, 12) - 6; small = rnd(e, 4); }   // y offset first (see generate_map)
                oy = coop_bcast(e, oy); ox = coop_bcast(e, ox); small = coop_bcast(e, small);
                if (small) gen_plop_s(e, x + ox, y + oy); else gen_plop_b(e, x + ox, y + oy);
                numPlops--;
            }
            numLakes--;
        }
    }
    gen_smooth_river(e);
    gen_trees(e);
    __syncwarp();
}
#endif

#undef S
} // namespace mp
