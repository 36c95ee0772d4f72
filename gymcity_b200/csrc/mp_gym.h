// mp_gym.h -- the gym layer of one env on the device: the TileMap shadow map with its
// clear-before-build action preprocessing (gym_city/envs/tilemap.py), the MicropolisControl
// window / tool tables (gym_city/envs/corecontrol.py) and MicropolisEnv's step / observation /
// reward (gym_city/envs/env.py).  Shared by mp_batch.cu and the CPU harness.
//
// The shadow map stores one zone id per window cell; the 22 one-hot feature planes the reference
// keeps (zoneMap[:-1]) are a pure function of that id (every write goes through
// zoneMap[:, x, y] = zoneCols[zone], tilemap.py:499-500), so they are expanded only when the
// observation is written.  Road-network labels (tilemap.py:201-330) are bookkeeping that never
// reaches the observation, reward or info and are not kept.
#pragma once
#include "mp_env.h"
#include "mp_gym_tables.h"

namespace mp {

enum : int { Z_POLICE = 5, Z_FIRE = 6, Z_AIRPORT = 7, Z_NUCLEAR = 8, Z_COAL = 9, Z_ROAD = 10, Z_RUBBLE = 14,
             Z_WATER = 16, Z_LAND = 17, Z_FOREST = 18 };
constexpr uint32_t F_PLANT = (1u << Z_NUCLEAR) | (1u << Z_COAL);
constexpr uint32_t F_ROAD = 1u << Z_ROAD;
constexpr uint32_t F_NATURAL = (1u << Z_LAND) | (1u << Z_WATER) | (1u << Z_RUBBLE) | (1u << Z_FOREST);
constexpr int GYM_MAX_TOOLS = 20, GYM_NUM_METRICS = 6, GYM_OBS_CHANNELS = 32;

struct GymCfg {
    int32_t W, H, xs, ys, num_tools, max_step;
    int8_t tool_zone[GYM_MAX_TOOLS];     // agent tool -> TileMap zone id; -1 = 'Nil', -2 = 'Clear'
    int8_t tool_engine[GYM_MAX_TOOLS];   // agent tool -> engine EditingTool
    double trg[GYM_NUM_METRICS], weight[GYM_NUM_METRICS];
    int32_t init_funds;
    // MicropolisEnv(poet=True) (env.py:301-313): populations divided by their parameter ranges and six more
    // scalar planes, the targets divided by theirs: 38 observation planes instead of 32
    int32_t poet;
    double param_range[GYM_NUM_METRICS];
};
MPH int gym_obs_channels(const GymCfg &c) { return c.poet ? GYM_OBS_CHANNELS + GYM_NUM_METRICS : GYM_OBS_CHANNELS; }

struct GymScalars {
    int32_t num_plants, num_roads, n_struct_tiles, total_traffic, num_step, tilemap_error;
    double metrics[GYM_NUM_METRICS], last_metrics[GYM_NUM_METRICS];
    double reward;
    int32_t done, curr_funds;
    int32_t track_ages;         // TileMap.track_ages (tilemap.py:201-206), set by the Extinguisher wrapper
    int32_t paint;              // TileMap.paint (tilemap.py:74-81): MicropolisPaintEnv's one-build-per-tile-per-step rule
    int32_t episode;            // resets this env has taken with derived seeds (gym_episode_seed): the episode index of the next one
};

// Per-episode seeds: hash of (run seed, GLOBAL env index, episode index, which) -> 31 bits; which = 0 terrain seed
// (generateSomeCity), 1 simRandom seed.  The same function as gymcity_b200/sharding.py mix(): a rollout does not
// depend on how the envs are sharded over GPUs, nor on whether the host or the device asked for the reset.
MPH int32_t gym_episode_seed(uint64_t seed, uint64_t global_env, uint64_t episode, uint64_t which)
{
    const uint64_t v[4] = { seed, global_env, episode, which };
    uint64_t h = 0x9e3779b97f4a7c15ull;
    for (int i = 0; i < 4; i++) {
        h ^= v[i] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        h *= 0xbf58476d1ce4e5b9ull;
        h ^= h >> 31;
    }
    return (int32_t)(h & 0x7fffffffull);
}

struct Gym {
    uint8_t *zone, *stat;       // [W*H] index x*H + y
    int16_t *center;            // [W*H] cx*128 + cy, -1 = None
    int32_t *age;               // [W*H] TileMap.age_order: num_step of the build, -1 = nothing built
    uint8_t *acted;             // [W*H] TileMap.acted: built on during this paint step
    GymScalars *g;
    const GymCfg *cfg;
};

MPH int gym_shadow_bytes(int W, int H)
{
    return a16(W * H) * 3 + a16(W * H * 2) + a16((int)sizeof(GymScalars)) + a16(W * H * 4);
}
MPH void bind_gym(Gym &gy, uint8_t *blk, const GymCfg *cfg)
{
    int n = cfg->W * cfg->H;
    gy.zone = blk;
    gy.stat = blk + a16(n);
    gy.center = (int16_t *)(blk + 2 * a16(n));
    gy.g = (GymScalars *)(blk + 2 * a16(n) + a16(n * 2));
    gy.age = (int32_t *)(blk + 2 * a16(n) + a16(n * 2) + a16((int)sizeof(GymScalars)));
    gy.acted = blk + 2 * a16(n) + a16(n * 2) + a16((int)sizeof(GymScalars)) + a16(n * 4);
    gy.cfg = cfg;
}

#define CFG (*gy.cfg)
#define G (*gy.g)

MPD int gym_cell(const Gym &gy, int x, int y) { return x * CFG.H + y; }
// MicropolisControl.getTile (corecontrol.py:302-305): engine.getTile(x + MAP_XS, y + MAP_YS) & 1023
MPD int gym_get_tile(const Env &e, const Gym &gy, int x, int y) { return tget(e, x + CFG.xs, y + CFG.ys) & 1023; }
MPD int gym_zone_of_tile(Gym &gy, int tile)
{
    MP_COV(gym_zone_of_tile);
    int z = kTileZone[tile];
    if (z == 255) {          // zoneSize[...] raises KeyError in the reference (Flood/Radioactive/Trash/Lightning)
        G.tilemap_error = 1;
        z = Z_RUBBLE;
    }
    return z;
}

// [LEAD] tilemap.py:467-541 updateTile.  zone < 0: re-read from the map.  static_build: -1 None, 0, 1.
MPN void gym_update_tile(Env &e, Gym &gy, int x, int y, int zone, int center, int static_build)
{
    MP_COV(gym_update_tile);
    const int c = gym_cell(gy, x, y);
    const uint32_t was = kZoneFeatures[gy.zone[c]];
    if (zone < 0) {
        zone = gym_zone_of_tile(gy, gym_get_tile(e, gy, x, y));
        center = (kZoneSize[zone] != 1) ? gy.center[c] : (x * 128 + y);
        if (static_build == 0) gy.stat[c] = 0;
    } else {
        gy.stat[c] = (static_build == 1 && zone != Z_LAND && zone != Z_RUBBLE && zone != Z_WATER) ? 1 : 0;
    }
    gy.zone[c] = (uint8_t)zone;
    gy.center[c] = (int16_t)center;
    const uint32_t is = kZoneFeatures[zone];
    if ((was & F_NATURAL) && !(is & F_NATURAL)) G.n_struct_tiles++;
    if (!(was & F_NATURAL) && (is & F_NATURAL)) G.n_struct_tiles--;
    if (G.track_ages && (is & F_NATURAL)) gy.age[c] = -1;          // tilemap.py:521-523
    if ((was & F_ROAD) && !(is & F_ROAD)) G.num_roads--;
    else if (!(was & F_ROAD) && (is & F_ROAD)) G.num_roads++;
    if ((was & F_PLANT) && !(is & F_PLANT)) G.num_plants--;
    else if (!(was & F_PLANT) && (is & F_PLANT)) G.num_plants++;
}

// [LEAD] tilemap.py:410-447 clearTile: Water, Land, Clear on the recorded centre, then re-read the
// old zone's footprint.
MPN void gym_clear_tile(Env &e, Gym &gy, int x, int y, bool static_build)
{
    MP_COV(gym_clear_tile);
    const int c = gym_cell(gy, x, y);
    const int old_zone = gy.zone[c];
    int cnt = gy.center[c];
    if (cnt < 0) cnt = x * 128 + y;
    if (gy.stat[c] == 1 && !static_build) return;
    int xc = cnt >> 7, yc = cnt & 127;
    // Water, Land, Clear on the centre.  Two thirds of the calls hit plain ground, where the three
    // tools collapse to (nearly) nothing, exactly:
    //  * bare DIRT (no status bits): layDoze fails without BULLBIT, Land sees DIRT -- three no-ops;
    //  * bulldozable rubble / trees with no road, rail or wire next to them (fixZone has nothing to
    //    redraw): Water dozes to DIRT for $1 and floods it, Land turns the RIVER back to DIRT, Clear
    //    is a no-op -- the tile ends as DIRT, $1 is spent, no RNG draw.
    // Everything else takes the three toolDowns.
    const int wx = xc + CFG.xs, wy = yc + CFG.ys;
    bool slow = true;
    if (inb(wx, wy)) {
        const int v = e.map()[tix(wx, wy)], t = v & LOMASK;
        if (v == T_DIRT) {
            slow = false;
        } else if ((v & BULLBIT) && !(v & ZONEBIT) && t >= T_TREEBASE && t <= T_LASTRUBBLE && e.s->totalFunds >= 1) {
            bool near = false;                              // a neighbour fixSingle could redraw
            MP_NOUNROLL for (int k = 0; k < 4; k++) {
                const int nx = wx + (k == 1) - (k == 3), ny = wy - (k == 0) + (k == 2);
                if (!inb(nx, ny)) continue;
                const int nt = e.map()[tix(nx, ny)] & LOMASK;
                near |= nt >= T_ROADBASE && nt < T_RESBASE;
            }
            if (!near) {
                mset(e, tix(wx, wy), T_DIRT);
                e.s->totalFunds -= 1;
                slow = false;
            }
        }
    }
    if (slow) {
        tool_down(e, TOOL_WATER, wx, wy);
        tool_down(e, TOOL_LAND, wx, wy);
        tool_down(e, TOOL_BULLDOZER, wx, wy);
    }
    const int size = kZoneSize[old_zone];
    if (size == 1) {
        if (xc < CFG.W && yc < CFG.H) gym_update_tile(e, gy, xc, yc, -1, -1, 0);
        return;
    }
    if (size == 6) { xc -= 1; yc -= 1; }
    int x0 = tmax(0, xc - 1), y0 = tmax(0, yc - 1);
    int x1 = tmin(xc - 1 + size, (int)CFG.W), y1 = tmin(yc - 1 + size, (int)CFG.H);
    MP_NOUNROLL for (int i = x0; i < x1; i++)
        MP_NOUNROLL for (int j = y0; j < y1; j++) gym_update_tile(e, gy, i, j, -1, -1, 0);
}

// [LEAD] tilemap.py:391-408 clearPatch
MPD int gym_clear_patch(Env &e, Gym &gy, int x, int y, int size, bool static_build)
{
    MP_COV(gym_clear_patch);
    if (size == 1 && !(!static_build && gy.stat[gym_cell(gy, x, y)] == 1)) {
        gym_clear_tile(e, gy, x, y, static_build);
        return 1;
    }
    int x0 = tmax(x - 1, 0), x1 = tmin(x - 1 + size, (int)CFG.W);
    int y0 = tmax(y - 1, 0), y1 = tmin(y - 1 + size, (int)CFG.H);
    MP_NOUNROLL for (int i = x0; i < x1; i++)
        MP_NOUNROLL for (int j = y0; j < y1; j++) {
            if (gy.stat[gym_cell(gy, i, j)] == 1) return 0;
            if (G.paint && gy.acted[gym_cell(gy, i, j)] == 1) return 0;     // built this turn: not overwritten (:402-404)
            gym_clear_tile(e, gy, i, j, static_build);
        }
    return 1;
}

// [LEAD] tilemap.py:450-479 addZone: the zone recorded is the one the MAP now shows at (x, y)
MPD void gym_add_zone(Env &e, Gym &gy, int x, int y, bool static_build)
{
    MP_COV(gym_add_zone);
    int zone = gym_zone_of_tile(gy, gym_get_tile(e, gy, x, y));
    int size = kZoneSize[zone];
    int center = (size == 6) ? (x + 1) * 128 + (y + 1) : x * 128 + y;
    if (size == 1) {
        gym_update_tile(e, gy, x, y, zone, center, static_build ? 1 : 0);
        if (G.paint) gy.acted[gym_cell(gy, x, y)] = 1;                   // :465-466
        if (G.track_ages) gy.age[gym_cell(gy, x, y)] = G.num_step;      // :467-468 (also after a 'Land' build)
        return;
    }
    int x0 = tmax(0, x - 1), y0 = tmax(0, y - 1);
    int x1 = tmin(x - 1 + size, (int)CFG.W), y1 = tmin(y - 1 + size, (int)CFG.H);
    MP_NOUNROLL for (int i = x0; i < x1; i++)
        MP_NOUNROLL for (int j = y0; j < y1; j++) {
            gym_update_tile(e, gy, i, j, zone, center, static_build ? 1 : 0);
            if (G.paint) gy.acted[gym_cell(gy, i, j)] = 1;               // :476-477
            if (G.track_ages) gy.age[gym_cell(gy, i, j)] = G.num_step;  // :478-479
        }
}

// The full agent tool list (corecontrol.py:85-106) as (TileMap zone id, engine EditingTool);
// zone -1 = 'Nil', -2 = 'Clear'.  Set-up builds (powerPuzzle, env.py:253-262) use tool NAMES that
// need not be in the env's own tool list, so they are addressed through this table.
#if defined(__CUDACC__)
#define MP_GYM_TABLE static __device__ const
#else
#define MP_GYM_TABLE static const
#endif
MP_GYM_TABLE int8_t kFullToolZone[GYM_MAX_TOOLS] = { 0, 1, 2, 6, 5, -2, 13, 11, 10, 4, 12, 3, 9, 8, 7, 15, 16, 17, 18, -1 };
MP_GYM_TABLE int8_t kFullToolEngine[GYM_MAX_TOOLS] = { 0, 1, 2, 3, 4, 7, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 0 };

// [LEAD] tilemap.py:345-377 addZoneBot for the tool recorded as `zone` and applied as `engine_tool`.
MPN bool gym_add_zone_bot(Env &e, Gym &gy, int x, int y, int zone, int engine_tool, bool static_build)
{
    MP_COV(gym_add_zone_bot);
    if (zone == -1) return true;                              // 'Nil'
    const int c = gym_cell(gy, x, y);
    const int old_zone = gy.zone[c];
    if (zone >= 0 && kZoneCompat[zone] != 0
        && (((kZoneCompat[zone] >> old_zone) & 1u) || ((kCompositeOf[old_zone] >> zone) & 1u)) && zone != Z_WATER) {
        if (gy.stat[c] == 1) static_build = true;
    }
    if (!static_build && gy.stat[c] == 1) return false;
    if (zone == -2) zone = Z_LAND;                            // 'Clear' is recorded as 'Land'
    int result;
    if (kZoneCompat[zone] != 0 && ((kZoneCompat[zone] >> old_zone) & 1u)) result = 1;   // clearForZone :382-389
    else result = gym_clear_patch(e, gy, x, y, kZoneSize[zone], static_build);
    if (result == 1) {
        result = tool_down(e, engine_tool, x + CFG.xs, y + CFG.ys);
        if (result == 1) gym_add_zone(e, gy, x, y, static_build);
    }
    return result == 1;
}

// [LEAD] one decoded action through addZoneBot; full != 0 decodes against the full agent tool list
MPD int gym_apply_action(Env &e, Gym &gy, long long a, bool static_build, bool full)
{
    MP_COV(gym_apply_action);
    if (a < 0) return 0;
    const int n = CFG.W * CFG.H;
    int z = (int)(a / n), rem = (int)(a - (long long)z * n);
    int x = rem / CFG.H, y = rem - x * CFG.H;
    if (full) {
        if (z >= GYM_MAX_TOOLS) return 0;
        return gym_add_zone_bot(e, gy, x, y, kFullToolZone[z], kFullToolEngine[z], static_build);
    }
    if (z >= CFG.num_tools) return 0;
    return gym_add_zone_bot(e, gy, x, y, CFG.tool_zone[z], CFG.tool_engine[z], static_build);
}

// ---- MicropolisPaintEnv (gym_city/envs/paintenv.py, paintcontrol.py) --------------------------
// [LEAD] MicropolisPaintControl.takeAction (paintcontrol.py:45-70): one tool per window cell, row by
// row, a cell is skipped if an earlier build of this step already covers it; the last column is never
// reached (the loop breaks at j >= MAP_Y - 1), as in the reference.  a: tool index per cell [W*H].
MPN void gym_paint_action(Env &e, Gym &gy, const uint8_t *a)
{
    MP_COV(gym_paint_action);
    const int W = CFG.W, H = CFG.H;
    G.paint = 1;                    // the acted rule applies to the builds of this action only
    MP_NOUNROLL for (int i = 0; i < W; i++)
        MP_NOUNROLL for (int j = 0; j < H - 1; j++) {
            const int c = gym_cell(gy, i, j), t = a[c];
            if (t >= CFG.num_tools || gy.acted[c]) continue;
            gym_add_zone_bot(e, gy, i, j, CFG.tool_zone[t], CFG.tool_engine[t], false);
        }
    MP_NOUNROLL for (int c = 0; c < W * H; c++) gy.acted[c] = 0;
    G.paint = 0;
}

// ---- Extinguisher wrapper (gym_city/wrappers.py:10-160) -----------------------------------------
// [ALL] TileMap.init_age_array (tilemap.py:201-206)
MPD void gym_init_age_array(Env &e, Gym &gy)
{
    MP_COV(gym_init_age_array);
    const int n = CFG.W * CFG.H;
    for (int i = e.c.tid; i < n; i += e.c.n) gy.age[i] = -1;
    if (e.c.lead()) G.track_ages = 1;
    e.c.sync();
}
// micro.doBotTool(x, y, 'Clear', static_build=True)
MPD void gym_extinct_clear(Env &e, Gym &gy, int x, int y) { gym_add_zone_bot(e, gy, x, y, -2, TOOL_BULLDOZER, true); }

enum : int { XT_AGE = 0, XT_SPATIAL = 1, XT_RANDOM = 2, XT_MONSTER = 3 };

// [LEAD] Extinguisher.extinguish (wrappers.py:42-55).  rnd: n_dels + 2 non-negative integers drawn by
// the host wrapper (the reference draws from numpy's global stream, which cannot be replayed):
// XT_RANDOM uses rnd[i] % (number of aged cells) for deletion i, XT_SPATIAL rnd[0] % W, rnd[1] % H as
// the centre.  Returns the number of deletions (curr_dels).
MPN int gym_extinguish(Env &e, Gym &gy, int type, int n_dels, const int32_t *rnd)
{
    MP_COV(gym_extinguish);
    const int W = CFG.W, H = CFG.H, n = W * H;
    int dels = 0;
    if (type == XT_MONSTER) { make_monster(e); return 0; }                       // engine.makeMonster()
    if (type == XT_AGE) {                                                        // elderCleanse :118-141
        MP_NOUNROLL for (int i = 0; i < n_dels; i++) {
            int youngest = -1, live = 0;
            MP_NOUNROLL for (int c = 0; c < n; c++) { const int a = gy.age[c]; youngest = tmax(youngest, a); live += a > -1; }
            if (!live) break;
            // ages += (ages < 0) * 2 * youngest; argmin (first minimum in flat order).  With youngest == 0
            // an unbuilt cell (-1) wins, as in the reference.
            int best = 0, bestv = 0x7fffffff;
            MP_NOUNROLL for (int c = 0; c < n; c++) {
                int a = gy.age[c];
                if (a < 0) a += 2 * youngest;
                if (a < bestv) { bestv = a; best = c; }
            }
            gym_extinct_clear(e, gy, best / H, best - (best / H) * H);
            dels++;
        }
        e.s->totalFunds = CFG.init_funds;                                       // :140
        return dels;
    }
    if (type == XT_RANDOM) {                                                     // ranDemolish :88-104
        MP_NOUNROLL for (int i = 0; i < n_dels; i++) {
            int live = 0;
            MP_NOUNROLL for (int c = 0; c < n; c++) live += gy.age[c] > -1;
            if (!live) break;
            int k = (int)((uint32_t)rnd[i] % (uint32_t)live), pick = -1;
            MP_NOUNROLL for (int c = 0; c < n; c++)
                if (gy.age[c] > -1 && k-- == 0) { pick = c; break; }
            gym_extinct_clear(e, gy, pick / H, pick - (pick / H) * H);
            dels++;
        }
        return dels;
    }
    // localWipe :57-67 + clear_border :69-86 (the y bound tests MAP_X, as in the reference)
    const int x = (int)((uint32_t)rnd[0] % (uint32_t)W), y = (int)((uint32_t)rnd[1] % (uint32_t)H);
    const int rmax = tmax(tmax(x, W - x < 0 ? x - W : W - x), tmax(y, H - y < 0 ? y - H : H - y));
    MP_NOUNROLL for (int r = 0; r < rmax; r++) {
        bool full = false;
        MP_NOUNROLL for (int xi = x - r; xi < x + r && !full; xi++) {
            if (xi < 0 || xi >= W) continue;
            MP_NOUNROLL for (int yi = y - r; yi < y + r; yi++) {
                if (yi < 0 || yi >= W || yi >= H) continue;
                if (gy.age[gym_cell(gy, xi, yi)] > 0) {
                    gym_extinct_clear(e, gy, xi, yi);
                    if (++dels == n_dels) { full = true; break; }
                }
            }
        }
        if (dels >= n_dels) break;
    }
    return dels;
}

// [ALL] TileMap.setEmpty (tilemap.py:327-343)
MPD void gym_set_empty(Env &e, Gym &gy)
{
    MP_COV(gym_set_empty);
    const int n = CFG.W * CFG.H;
    for (int i = e.c.tid; i < n; i += e.c.n) { gy.zone[i] = Z_LAND; gy.stat[i] = 0; gy.center[i] = -1; }
    if (e.c.lead()) { G.num_roads = 0; G.num_plants = 0; }
    e.c.sync();
}

// [ALL] MicropolisControl.updateMap (corecontrol.py:193-203): every window cell re-read with
// zone given, centre (i, j), static_build None
MPD void gym_update_map(Env &e, Gym &gy)
{
    MP_COV(gym_update_map);
    if (e.c.lead())
        MP_NOUNROLL for (int i = 0; i < CFG.W; i++)
            MP_NOUNROLL for (int j = 0; j < CFG.H; j++)
                gym_update_tile(e, gy, i, j, gym_zone_of_tile(gy, gym_get_tile(e, gy, i, j)), i * 128 + j, -1);
    e.c.sync();
}

// [LEAD] env.py:426-440 get_city_metrics
MPD void gym_city_metrics(const Env &e, Gym &gy, double *m)
{
    MP_COV(gym_city_metrics);
    m[0] = e.s->resPop; m[1] = e.s->comPop; m[2] = e.s->indPop;
    m[3] = G.total_traffic; m[4] = G.num_plants; m[5] = e.s->cityYes;
}

MPD int dsign(double v) { return (v > 0) - (v < 0); }
MPD double dabs(double v) { return v < 0 ? -v : v; }

// [LEAD] env.py:345-387 getReward
MPD double gym_reward(const Gym &gy)
{
    MP_COV(gym_reward);
    double reward = 0;
    MP_NOUNROLL for (int k = 0; k < GYM_NUM_METRICS; k++) {
        double last = G.last_metrics[k], trg_change = CFG.trg[k] - last, change = G.metrics[k] - last, r;
        if (dsign(change) != dsign(trg_change)) r = -dabs(change);
        else if (dabs(change) < dabs(trg_change)) r = dabs(change);
        else r = dabs(trg_change) - dabs(trg_change - change);
        reward += r * CFG.weight[k];
    }
    return reward;
}

// [ALL] env.py:301-343 getState/observation + corecontrol.py:215-239 getDensityMaps: writes the
// (32, W, H) float32 observation and refreshes total_traffic.  Channel order: SURVEY appendix A.3.
MPD void gym_observe(Env &e, Gym &gy, float *obs)
{
    MP_COV(gym_observe);
    const int W = CFG.W, H = CFG.H, n = W * H;
    if (e.c.lead()) e.c.red[3] = 0;
    e.c.sync();
    const Scalars &s = *e.s;
    float sc[6 + GYM_NUM_METRICS];
    sc[0] = (float)s.resPop; sc[1] = (float)s.comPop; sc[2] = (float)s.indPop;
    sc[3] = (float)s.resValve / (float)RES_VALVE_RANGE;        // utilities.cpp:416-424 getDemands
    sc[4] = (float)s.comValve / (float)COM_VALVE_RANGE;
    sc[5] = (float)s.indValve / (float)IND_VALVE_RANGE;
    const int nsc = CFG.poet ? 6 + GYM_NUM_METRICS : 6;
    if (CFG.poet) {                                            // env.py:306-312 (Python floats: double division)
        sc[0] = (float)((double)s.resPop / CFG.param_range[0]);
        sc[1] = (float)((double)s.comPop / CFG.param_range[1]);
        sc[2] = (float)((double)s.indPop / CFG.param_range[2]);
        MP_NOUNROLL for (int k = 0; k < GYM_NUM_METRICS; k++) sc[6 + k] = (float)(CFG.trg[k] / CFG.param_range[k]);
    }
    int tsum = 0;
    // per-cell inputs of plane 22 (power), 23 (population density), 24 (traffic density); returns the traffic value the
    // cell contributes to total_traffic (each half-resolution cluster once)
    auto cell_inputs = [&](int cell, float &pw, float &vp, float &vt) -> int {
        const int i = cell / H, j = cell - i * H;
        pw = (float)pwr_get(e, j + CFG.xs, i + CFG.ys);
        // cluster (x = j' + MAP_XS/2, y = i' + MAP_YS/2) fills rows 2i'..2i'+1, cols 2j'..2j'+1
        vp = 0.0f; vt = 0.0f;
        const int i2 = i >> 1, j2 = j >> 1;
        if (i2 < W / 2 && j2 < H / 2) {
            const int cx = j2 + CFG.xs / 2, cy = i2 + CFG.ys / 2;
            if (cx >= 0 && cx < W2 && cy >= 0 && cy < H2) {
                const int dp = e.pop()[cx * H2 + cy], dt = e.traffic()[cx * H2 + cy];
                vp = (float)dp;
                vt = (float)dt;
                if (!(i & 1) && !(j & 1)) return dt;
            }
        }
        return 0;
    };
#if defined(__CUDA_ARCH__)
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0) {
        // four consecutive window cells per lane and iteration: every plane is written with 128-bit stores, 512
        // contiguous bytes per warp instruction (a 120x100 window is 1.5 MB of observation per env: with 32-bit
        // stores this kernel ran at a quarter of the HBM rate, profiles/r02_notes.md)
        for (int c0 = e.c.tid * 4; c0 < n; c0 += e.c.n * 4) {
            const uint32_t z4 = *reinterpret_cast<const uint32_t *>(gy.zone + c0);
            const uint32_t s4 = *reinterpret_cast<const uint32_t *>(gy.stat + c0);
            uint32_t feat[4];
            float pw[4], vp[4], vt[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                feat[q] = kZoneFeatures[(z4 >> (8 * q)) & 0xffu];
                tsum += cell_inputs(c0 + q, pw[q], vp[q], vt[q]);
            }
            float4 *o = reinterpret_cast<float4 *>(obs + c0);
            const int n4 = n >> 2;
            MP_NOUNROLL for (int p = 0; p < kNumFeatures; p++)
                o[p * n4] = make_float4((float)((feat[0] >> p) & 1u), (float)((feat[1] >> p) & 1u),
                                        (float)((feat[2] >> p) & 1u), (float)((feat[3] >> p) & 1u));
            o[22 * n4] = make_float4(pw[0], pw[1], pw[2], pw[3]);
            o[23 * n4] = make_float4(vp[0], vp[1], vp[2], vp[3]);
            o[24 * n4] = make_float4(vt[0], vt[1], vt[2], vt[3]);
            MP_NOUNROLL for (int k = 0; k < nsc; k++) o[(25 + k) * n4] = make_float4(sc[k], sc[k], sc[k], sc[k]);
            o[(25 + nsc) * n4] = make_float4((float)(s4 & 0xffu), (float)((s4 >> 8) & 0xffu), (float)((s4 >> 16) & 0xffu),
                                             (float)((s4 >> 24) & 0xffu));
        }
    } else
#endif
    // one window cell per lane and iteration: every plane's store is coalesced across the lanes, and
    // the per-cell inputs (zone id, power bit, the two half-resolution densities) are read once
    for (int cell = e.c.tid; cell < n; cell += e.c.n) {
        float *o = obs + cell;
        const uint32_t feat = kZoneFeatures[gy.zone[cell]];
        MP_NOUNROLL for (int p = 0; p < kNumFeatures; p++) o[p * n] = (float)((feat >> p) & 1u);
        float pw, vp, vt;
        tsum += cell_inputs(cell, pw, vp, vt);
        o[22 * n] = pw;
        o[23 * n] = vp;
        o[24 * n] = vt;
        MP_NOUNROLL for (int k = 0; k < nsc; k++) o[(25 + k) * n] = sc[k];
        o[(25 + nsc) * n] = (float)gy.stat[cell];
    }
    coop_add(e, 3, tsum);
    e.c.sync();
    if (e.c.lead()) G.total_traffic = e.c.red[3];
    e.c.sync();
}

// [LEAD] MicropolisEnv.step part 1 (env.py:448-466): takeAction + micro.setFunds(init_funds).
// action < 0: no agent action.
MPD void gym_step_pre(Env &e, Gym &gy, long long action, bool static_build)
{
    MP_COV(gym_step_pre);
    gym_apply_action(e, gy, action, static_build, false);       // takeAction
    e.s->totalFunds = CFG.init_funds;                           // micro.setFunds(init_funds)
}

// [ALL] MicropolisEnv.postact after micro.simTick() (env.py:468-540): state, metrics, reward, done
MPD void gym_step_post(Env &e, Gym &gy, float *obs, float *reward_out, uint8_t *done_out)
{
    MP_COV(gym_step_post);
    gym_observe(e, gy, obs);                                    // getState()
    if (e.c.lead()) {
        MP_NOUNROLL for (int k = 0; k < GYM_NUM_METRICS; k++) G.last_metrics[k] = G.metrics[k];
        gym_city_metrics(e, gy, G.metrics);
        G.reward = gym_reward(gy);
        G.curr_funds = (int32_t)e.s->totalFunds;
        G.done = (e.s->totalFunds < 0 || G.num_step >= CFG.max_step) ? 1 : 0;
        G.num_step++;
        *reward_out = (float)G.reward;
        *done_out = (uint8_t)G.done;
    }
    e.c.sync();
}

// [ALL] the whole step for one env (CPU harness; the CUDA path splits it into launches)
MPD void gym_step(Env &e, Gym &gy, long long action, bool static_build, float *obs, float *reward_out,
                  uint8_t *done_out)
{
    MP_COV(gym_step);
    if (e.c.lead()) gym_step_pre(e, gy, action, static_build);
    e.c.sync();
    control_sim_tick(e);                                        // micro.simTick()
    gym_step_post(e, gy, obs, reward_out, done_out);
}

// [ALL] MicropolisEnv.reset, first half (env.py:264-274): clearMap (+ newMap) and num_step = 0.
// terrain: 0 = empty_start (clearMap only), 1 = generateSomeCity(terrain_seed) + seedRandom(sim_seed).
MPD void gym_reset_begin(Env &e, Gym &gy, int terrain, int terrain_seed, int sim_seed, const uint16_t *prepared = nullptr)
{
    MP_COV(gym_reset_begin);
    gym_set_empty(e, gy);
    clear_map(e);
    gym_update_map(e, gy);
    if (terrain) {
        generate_some_city(e, terrain_seed, sim_seed, prepared);
        gym_update_map(e, gy);
    }
    if (e.c.lead()) G.num_step = 0;
    e.c.sync();
}

// [ALL] MicropolisEnv.reset after micro.simTick() (env.py:280-292): metrics, funds, reward, state
// keep_outputs: the worker's auto-reset (subproc_vec_env.py:16-21) hands back the reset observation together with the
// reward and done flag of the step that ended the episode
MPD void gym_reset_post(Env &e, Gym &gy, float *obs, float *reward_out, uint8_t *done_out, bool keep_outputs = false)
{
    MP_COV(gym_reset_post);
    if (e.c.lead()) {
        gym_city_metrics(e, gy, G.metrics);                     // uses the previous total_traffic
        MP_NOUNROLL for (int k = 0; k < GYM_NUM_METRICS; k++) G.last_metrics[k] = G.metrics[k];
        e.s->totalFunds = CFG.init_funds;
        G.reward = gym_reward(gy);
        G.done = 0;
        if (!keep_outputs) {
            *reward_out = (float)G.reward;
            *done_out = 0;
        }
    }
    e.c.sync();
    gym_observe(e, gy, obs);
}

// [ALL] MicropolisEnv.reset, second half (env.py:279-292) for one env (CPU harness)
MPD void gym_reset_finish(Env &e, Gym &gy, float *obs, float *reward_out, uint8_t *done_out)
{
    MP_COV(gym_reset_finish);
    control_sim_tick(e);
    gym_reset_post(e, gy, obs, reward_out, done_out);
}

#undef CFG
#undef G
} // namespace mp
