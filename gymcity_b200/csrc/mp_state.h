// mp_state.h -- per-city state layout, accessors, RNG and the cooperative-thread context.
//
// One city ("env") is one contiguous block in HBM (see EnvLayout): the 120x100 uint16 tile map
// (column-major x*100+y like the reference, allocate.cpp:75-87), the five 60x50 byte overlays,
// the 30x25 terrain overlay, the power grid BIT-PACKED (12 000 bits), the six 15x13 short maps,
// the history arrays, the scalar block and the sprite table, followed by per-env scratch.
// The engine code in mp_*.h is written once against `Env` (a bundle of pointers + a thread
// context) and compiles both for the device (one CTA per env, `Coop` = the CTA) and, for the CPU
// test harness under tests/, for the host (`Coop` = a single thread).
#pragma once
#include <stdint.h>
#include <string.h>
#include "mp_tiles.h"

#if defined(__CUDACC__)
#define MPD __device__ __forceinline__
#define MPN __device__ __noinline__
#define MPH __host__ __device__ inline
#define MPA __host__ __device__ __forceinline__
// MP_TAB: small lookup tables of the engine.  Function-scope `static const`: one copy in constant memory, read with an
// indexed load -- as a plain local array the compiler rebuilds the table on the stack at every call (a few STL) and
// reads it back through local memory on the serial lane's dependency chain.
#define MP_TAB static const
// The serial lane is latency bound and its code footprint is what hurts (the engine is ~10x the SM
// instruction caches): keep its loops rolled.
#define MP_NOUNROLL _Pragma("unroll 1")
// MPP: inlined in the product build, out of line when profiling (-DMP_PROFILE_SPLIT) so that ncu's
// SASS samples can be attributed per engine function (tools/ncu_by_function.py).
#if defined(MP_PROFILE_SPLIT)
#define MPP __device__ __noinline__
#else
#define MPP __device__ __forceinline__
#endif
// MPQ: the per-tile dispatch targets (doZone, doRoad, doRail) and the per-frame entry points (sim_phase, moveObjects).
// Inline: a call on the serial lane is a jump into code the SM has not got -- one call per 200 instructions was worth
// 9 % of the step (profiles/r02_notes.md); out of line only for per-function profiles.
#if defined(MP_PROFILE_SPLIT)
#define MPQ __device__ __noinline__
#else
#define MPQ __device__ __forceinline__
#endif
#define MPR MPQ
#define MPS MPQ         // the sprite handlers moveObjects dispatches to, destroyMapTile
#else
// host harness: the 128-bit vector type the device code moves cells with
struct uint4 { unsigned int x, y, z, w; };
inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { uint4 v = { x, y, z, w }; return v; }
#define MPQ inline
#define MP_TAB static const
#define MPR inline
#define MPS inline
#define MPP inline
#define MPD inline
#define MPN inline
#define MPH inline
#define MPA inline
#define MP_NOUNROLL
#endif

// MP_WARP_COOP: the code is compiled for the device with one WARP per env (lanes cooperate through shuffles,
// ballots and warp barriers).  Otherwise one THREAD runs an env on its own: the host harness under tests/.
// (A thread-per-env DEVICE build, 32 different envs in the lanes of a warp, is bit-exact too and 2.8x slower:
// profiles/r02_notes.md.)
#if defined(__CUDA_ARCH__)
#define MP_WARP_COOP 1
#else
#define MP_WARP_COOP 0
#endif

#include "mp_cov.h"

namespace mp {

constexpr int WORLD_W = 120, WORLD_H = 100, NTILES = WORLD_W * WORLD_H;
constexpr int W2 = 60, H2 = 50, N2 = W2 * H2;       // MapByte2
constexpr int W4 = 30, H4 = 25, N4 = W4 * H4;       // MapByte4
constexpr int W8 = 15, H8 = 13, N8 = W8 * H8;       // MapShort8
constexpr int POWER_WORDS = NTILES / 32;            // 375
constexpr int POWER_STACK_SIZE = NTILES / 4;        // micropolis.h:221
constexpr int MAX_TRAFFIC_DISTANCE = 30;
constexpr int MAX_ROAD_EFFECT = 32, MAX_POLICE_EFFECT = 1000, MAX_FIRE_EFFECT = 1000;
constexpr int RES_VALVE_RANGE = 2000, COM_VALVE_RANGE = 1500, IND_VALVE_RANGE = 1500;
constexpr int PROBNUM = 10, NUMPROBLEMS = 7, PROBLEM_COMPLAINTS = 4;
constexpr int MAX_SPRITES = 48;

enum : int {
    PWRBIT = 0x8000, CONDBIT = 0x4000, BURNBIT = 0x2000, BULLBIT = 0x1000, ANIMBIT = 0x0800,
    ZONEBIT = 0x0400, ALLBITS = 0xfc00, LOMASK = 0x03ff,
    BLBNBIT = BULLBIT | BURNBIT, BLBNCNBIT = BULLBIT | BURNBIT | CONDBIT, BNCNBIT = BURNBIT | CONDBIT
};

enum SpriteType : int {
    SPR_NONE = 0, SPR_TRAIN, SPR_HELICOPTER, SPR_AIRPLANE, SPR_SHIP, SPR_MONSTER, SPR_TORNADO,
    SPR_EXPLOSION, SPR_BUS, SPR_COUNT
};

enum Dir : int { D_INVALID = 0, D_N, D_NE, D_E, D_SE, D_S, D_SW, D_W, D_NW, D_END };

// reference micropolis.h:981-1018 (SimSprite); list links are slot indices, -1 = end
struct Sprite {
    int16_t type, frame, x, y, width, height, xOffset, yOffset, xHot, yHot;
    int16_t origX, origY, destX, destY, count, soundCount, dir, newDir;
    int16_t step, flag, control, turn, accel, speed;
    int16_t next, used;
};

// Scalar state; widths follow reference micropolis.h (short / int / Quad=long / float / bool).
struct Scalars {
    int64_t cityTime, roadSpend, policeSpend, fireSpend, roadFund, policeFund, fireFund;
    int64_t roadEffect, policeEffect, fireEffect, taxFund, roadValue, policeValue, fireValue;
    int64_t cityPop, cityPopDelta, cityAssessedValue, cityPopLast, totalFunds;
    uint32_t rng;
    int32_t powerStackPointer, absDist, simPasses, gameLevel, cityClass, disasterEvent, scoreType,
        scenario;
    float roadPercent, policePercent, firePercent, externalMarket;
    int16_t roadTotal, railTotal, firePop, resPop, comPop, indPop, totalPop, totalPopLast;
    int16_t resZonePop, comZonePop, indZonePop, totalZonePop, hospitalPop, churchPop, faith;
    int16_t stadiumPop, policeStationPop, fireStationPop, coalPowerPop, nuclearPowerPop;
    int16_t seaportPop, airportPop, crimeAverage, pollutionAverage, landValueAverage, cityTax;
    int16_t needHospital, needChurch, floodCount, cityYes, cityScore, cityScoreDelta;
    int16_t trafficAverage, categoryLast, cityCenterX, cityCenterY, pollutionMaxX, pollutionMaxY;
    int16_t crimeMaxX, crimeMaxY, crimeRamp, pollutionRamp, cashFlow, disasterWait, scoreWait;
    int16_t poweredZoneCount, unpoweredZoneCount, cityTaxAverage, simCycle, phaseCycle, speedCycle;
    int16_t resValve, comValve, indValve, spriteCycle, initSimLoad, simSpeed, trafMaxX, trafMaxY;
    int16_t problemVotes[PROBNUM], problemOrder[PROBLEM_COMPLAINTS];
    int16_t spriteHead, globalSprites[SPR_COUNT], curMapStackPointer;
    uint8_t taxFlag, resCap, comCap, indCap, newPower, doInitialEval, autoBulldoze, autoBudget;
    uint8_t enableDisasters, doAnimation, tilesAnimated, spriteOverflow;
    // derived, not part of the reference state -- memo for doPowerScan.  The grid a scan produces is a
    // function of the map's conductive layout and of the seeds pushed since the previous scan (plant
    // centres, including stale ones of plants removed meanwhile).  powerQuiet: no write since the last scan has
    // flipped the CONDBIT of a tile that is powered, touches a powered tile or is a plant centre (mset) and no
    // bulk map / state write happened.  powerPure: the
    // last scan followed a quiet interval (all its seeds were current plants) and stayed clear of the
    // capacity cut-off.  A scan that finds both set would recompute the grid it already has.
    uint8_t powerQuiet, powerPure, pad_[6];
    int16_t curMapStack[MAX_TRAFFIC_DISTANCE + 2];   // packed x*100+y, entries 1..pointer
};
static_assert(sizeof(Scalars) % 8 == 0, "Scalars is copied to shared memory word by word");

// Byte offsets of one env's block.  Every section starts 16-byte aligned so it can be moved
// with 128-bit loads / bulk copies.
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int a16(int v) { return (v + 15) & ~15; }
struct EnvLayout {
    static constexpr int MAP = 0;                                   // uint16[12000]
    static constexpr int POP = MAP + NTILES * 2;                    // uint8[3000] x5
    static constexpr int TRAFFIC = POP + a16(N2);
    static constexpr int POLLUTION = TRAFFIC + a16(N2);
    static constexpr int LANDVAL = POLLUTION + a16(N2);
    static constexpr int CRIME = LANDVAL + a16(N2);
    static constexpr int TERRAIN = CRIME + a16(N2);                 // uint8[750]
    static constexpr int POWER = TERRAIN + a16(N4);                 // uint32[375]
    static constexpr int ROG = POWER + a16(POWER_WORDS * 4);        // int16[195] x6
    static constexpr int FIREST = ROG + a16(N8 * 2);
    static constexpr int FIREEFF = FIREST + a16(N8 * 2);
    static constexpr int POLICEST = FIREEFF + a16(N8 * 2);
    static constexpr int POLICEEFF = POLICEST + a16(N8 * 2);
    static constexpr int COMRATE = POLICEEFF + a16(N8 * 2);
    static constexpr int HIST = COMRATE + a16(N8 * 2);              // int16[6*240 + 120]
    static constexpr int SCALARS = HIST + a16((6 * 240 + 120) * 2);
    static constexpr int SPRITES = SCALARS + a16((int)sizeof(Scalars));
    static constexpr int PERSISTENT_END = SPRITES + a16(MAX_SPRITES * (int)sizeof(Sprite));
    // scratch (not part of the persistent state)
    static constexpr int TMP1 = PERSISTENT_END;                     // uint8[3000]
    static constexpr int TMP2 = TMP1 + a16(N2);
    static constexpr int TMP3 = TMP2 + a16(N2);                     // uint8[750]
    static constexpr int STMP = TMP3 + a16(N4);                     // int16[195]
    static constexpr int PSTACK = STMP + a16(N8 * 2);               // uint16[3000]
    static constexpr int OVERLAY = PSTACK + a16(POWER_STACK_SIZE * 2);  // tool overlay uint16[256]+mask
    // derived: one bit per tile, set iff mapScan would not skip it (tile id >= FLOOD)
    static constexpr int ACTIVE = OVERLAY + 512 + 32;
    // derived: subset of ACTIVE the scan really has to visit (see tile_needs_visit)
    static constexpr int NEED = ACTIVE + a16((POWER_WORDS + 1) * 4);
    static constexpr int TOTAL = NEED + a16((POWER_WORDS + 1) * 4);
};

// Thread context.  Device: the WARP that owns the env (tid = lane, n = 32; a barrier is a
// __syncwarp, so the warps of a CTA -- one env each -- never wait on one another).  Host harness:
// one thread.  Nothing in it (or in Env) differs between the lanes of a warp -- `tid` reads the
// lane id from the hardware -- so one Env per warp lives in SHARED memory instead of one copy per
// thread in local memory (30 pointers x 64 threads x 16 CTAs would not even fit the L1).
struct LaneId {
    MPD operator int() const
    {
#if MP_WARP_COOP
        return (int)(threadIdx.x & 31u);
#else
        return 0;
#endif
    }
};
struct Coop {
    LaneId tid;
#if defined(__CUDACC__)
    static constexpr int n = 32;
#else
    static constexpr int n = 1;
#endif
    int *red;       // >= 8 ints of scratch per env: shared memory (warp per env) / plain memory (thread per env, host)
    MPD void sync() const
    {
#if MP_WARP_COOP
        __syncwarp();
#endif
    }
    MPD bool lead() const { return (int)tid == 0; }
};

// Scalars pointer.  On the device every kernel stages the env's scalar block in SHARED memory for the
// duration of the launch (census counters, cycle counters, RNG state are touched on almost every serial
// instruction).  The copies live in the kernel's dynamic shared memory and the pointer is kept as a byte
// offset from that symbol, so the compiler knows the address space and emits LDS / STS with 32-bit addresses
// instead of generic-space accesses.  Host code (snapshots, the CPU harness) holds a plain pointer.
#if defined(__CUDACC__)
extern __shared__ __align__(16) uint8_t mp_dyn_smem[];
#endif
struct ScalPtr {
    Scalars *p;                 // device: byte offset into mp_dyn_smem
    MPA Scalars *get() const
    {
#if defined(__CUDA_ARCH__)
        return reinterpret_cast<Scalars *>(mp_dyn_smem + (uint32_t)reinterpret_cast<uintptr_t>(p));
#else
        return p;
#endif
    }
    MPA Scalars *operator->() const { return get(); }
    MPA Scalars &operator*() const { return *get(); }
    MPA operator Scalars *() const { return get(); }
    MPA ScalPtr &operator=(Scalars *q)
    {
#if defined(__CUDA_ARCH__)
        p = reinterpret_cast<Scalars *>((uintptr_t)(reinterpret_cast<uint8_t *>(q) - mp_dyn_smem));
#else
        p = q;
#endif
        return *this;
    }
};

// One env = the base pointer of its block; every map is the base plus a compile-time offset (EnvLayout), so
// a field access is one global load with an immediate offset -- no per-field pointer to fetch first.
struct Env {
    uint8_t *blk;
    ScalPtr s;
    uint16_t *ovl;          // tool overlay: 256 values + 8 mask words (block scratch, or shared memory in k_mp_env_pre)
    const uint16_t *anim;   // 1024-entry animation successor table
    Coop c;
    MPA uint8_t *base() const
    {
        uint8_t *b = blk;
#if defined(__CUDA_ARCH__) && !defined(MP_NO_ASSUME_GLOBAL)
        __builtin_assume(__isGlobal(b));
#endif
        return b;
    }
    typedef EnvLayout L;
    MPA uint16_t *map() const { return (uint16_t *)(base() + L::MAP); }
    MPA uint8_t *pop() const { return base() + L::POP; }
    MPA uint8_t *traffic() const { return base() + L::TRAFFIC; }
    MPA uint8_t *pollution() const { return base() + L::POLLUTION; }
    MPA uint8_t *landval() const { return base() + L::LANDVAL; }
    MPA uint8_t *crime() const { return base() + L::CRIME; }
    MPA uint8_t *terrain() const { return base() + L::TERRAIN; }
    MPA uint32_t *power() const { return (uint32_t *)(base() + L::POWER); }
    MPA int16_t *rog() const { return (int16_t *)(base() + L::ROG); }
    MPA int16_t *fireSt() const { return (int16_t *)(base() + L::FIREST); }
    MPA int16_t *fireEff() const { return (int16_t *)(base() + L::FIREEFF); }
    MPA int16_t *policeSt() const { return (int16_t *)(base() + L::POLICEST); }
    MPA int16_t *policeEff() const { return (int16_t *)(base() + L::POLICEEFF); }
    MPA int16_t *comRate() const { return (int16_t *)(base() + L::COMRATE); }
    MPA int16_t *resHist() const { return (int16_t *)(base() + L::HIST); }
    MPA int16_t *comHist() const { return resHist() + 240; }
    MPA int16_t *indHist() const { return resHist() + 480; }
    MPA int16_t *moneyHist() const { return resHist() + 720; }
    MPA int16_t *pollutionHist() const { return resHist() + 960; }
    MPA int16_t *crimeHist() const { return resHist() + 1200; }
    MPA int16_t *miscHist() const { return resHist() + 1440; }
    MPA Sprite *spr() const { return (Sprite *)(base() + L::SPRITES); }
    MPA uint8_t *tmp1() const { return base() + L::TMP1; }
    MPA uint8_t *tmp2() const { return base() + L::TMP2; }
    MPA uint8_t *tmp3() const { return base() + L::TMP3; }
    MPA int16_t *stmp() const { return (int16_t *)(base() + L::STMP); }
    MPA uint16_t *pstack() const { return (uint16_t *)(base() + L::PSTACK); }
    MPA uint32_t *act() const { return (uint32_t *)(base() + L::ACTIVE); }      // active-tile bitmap (derived from map, kept in step by mset)
    MPA uint32_t *need() const { return (uint32_t *)(base() + L::NEED); }       // tiles mapScan must visit: act minus building tiles whose PWRBIT is already right
};

// Scalars live in the block; the kernels re-point `s` at their shared-memory copy (stage_scalars).
MPH void bind_env(Env &e, uint8_t *blk)
{
    e.blk = blk;
#if !defined(__CUDA_ARCH__)
    e.s.p = (Scalars *)(blk + EnvLayout::SCALARS);      // device: the kernels point `s` at their shared-memory copy
#endif
    e.ovl = (uint16_t *)(blk + EnvLayout::OVERLAY);
}

// ------------------------------------------------------------------ small helpers
template <typename T> MPD T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> MPD T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> MPD T tclamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
MPD int iabs(int v) { return v < 0 ? -v : v; }
// float -> short the way x86-64 does it: truncate to int32, keep the low 16 bits
MPD int16_t f2s(float f) { return (int16_t)(int)f; }
MPD int16_t d2s(double f) { return (int16_t)(int)f; }

MPD bool inb(int x, int y) { return x >= 0 && x < WORLD_W && y >= 0 && y < WORLD_H; }
MPD int tix(int x, int y) { return x * WORLD_H + y; }

// Map<Byte,2>/<Byte,4>/<short,8>::worldGet / worldSet (map_type.h:225-300): off-map reads 0,
// off-map writes are dropped.
MPD int b2get(const uint8_t *m, int wx, int wy) { return inb(wx, wy) ? m[(wx >> 1) * H2 + (wy >> 1)] : 0; }
MPD void b2set(uint8_t *m, int wx, int wy, int v) { if (inb(wx, wy)) m[(wx >> 1) * H2 + (wy >> 1)] = (uint8_t)v; }
MPD int s8get(const int16_t *m, int wx, int wy) { return inb(wx, wy) ? m[(wx >> 3) * H8 + (wy >> 3)] : 0; }
MPD void s8set(int16_t *m, int wx, int wy, int v) { if (inb(wx, wy)) m[(wx >> 3) * H8 + (wy >> 3)] = (int16_t)v; }
MPD int pwr_get(const Env &e, int x, int y)
{
    if (!inb(x, y)) return 0;
    int i = tix(x, y);
    return (e.power()[i >> 5] >> (i & 31)) & 1u;
}
MPD void pwr_set(Env &e, int x, int y)
{
    if (!inb(x, y)) return;
    int i = tix(x, y);
    e.power()[i >> 5] |= 1u << (i & 31);
}
// Every tile store on the serial lane goes through mset so that the two derived bitmaps never go
// stale.  `act`: bit set iff tile id >= FLOOD, i.e. the reference's mapScan would look at the tile
// (simulate.cpp:937-945).  `need` (what the scan iterates): act minus the tiles whose visit is
// provably a no-op -- building tiles that are not a zone centre, road, rail or tiny explosion do
// nothing in mapScan except `if (newPower && CONDBIT) setZonePower` (simulate.cpp:966-968), which
// changes nothing when PWRBIT already agrees with the power grid.  No draw, no census: skipping
// such a visit is exact.  The bit comes back whenever the tile is rewritten (mset) or the power
// grid is rebuilt (refresh_need after doPowerScan).
MPH bool tile_is_active(int v) { return (v & LOMASK) >= T_FLOOD; }
MPD bool tile_is_inert_kind(int v)
{
    const int t = v & LOMASK;
    return t >= T_POWERBASE && !(v & ZONEBIT) && !(t >= T_RAILBASE && t < T_RESBASE)
        && !(t >= T_SOMETINYEXP && t <= T_LASTTINYEXP);
}
MPD bool tile_needs_visit(const Env &e, int idx, int v)
{
    if (!tile_is_active(v)) return false;
    if (!tile_is_inert_kind(v)) return true;
    if (!(v & CONDBIT) || !e.s->newPower) return false;
    const int t = v & LOMASK;
    const bool want = t == T_NUCLEAR || t == T_POWERPLANT || ((e.power()[idx >> 5] >> (idx & 31)) & 1u);
    return ((v & PWRBIT) != 0) != want;
}
#if defined(MP_MSET_NOINLINE)
MPN void mset(Env &e, int idx, int v)
#else
MPD void mset(Env &e, int idx, int v)
#endif
{
    const int old = e.map()[idx];
    if (old == (int)(uint16_t)v) return;                 // repairZone, smoke and traffic art mostly rewrite what is there
    if ((old ^ v) & CONDBIT) {
        // Does the flip touch the powered grid?  A tile that stops conducting matters if it was powered; a
        // tile that starts conducting matters if it is a plant centre (a new seed) or touches a powered tile
        // (anything joined to the grid later is joined through such a tile, whose own write is then caught).
        bool matters;
        if (v & CONDBIT) {
            const int t = v & LOMASK, x = idx / WORLD_H, y = idx - x * WORLD_H;
            matters = t == T_POWERPLANT || t == T_NUCLEAR;
            if (y > 0) matters |= ((e.power()[(idx - 1) >> 5] >> ((idx - 1) & 31)) & 1u) != 0;
            if (y < WORLD_H - 1) matters |= ((e.power()[(idx + 1) >> 5] >> ((idx + 1) & 31)) & 1u) != 0;
            if (x > 0) matters |= ((e.power()[(idx - WORLD_H) >> 5] >> ((idx - WORLD_H) & 31)) & 1u) != 0;
            if (x < WORLD_W - 1) matters |= ((e.power()[(idx + WORLD_H) >> 5] >> ((idx + WORLD_H) & 31)) & 1u) != 0;
        } else {
            matters = ((e.power()[idx >> 5] >> (idx & 31)) & 1u) != 0;
        }
        if (matters) e.s->powerQuiet = 0;
    }
    e.map()[idx] = (uint16_t)v;
    const uint32_t bit = 1u << (idx & 31);
    const uint32_t w = e.act()[idx >> 5], nw = tile_is_active(v) ? (w | bit) : (w & ~bit);
    if (nw != w) e.act()[idx >> 5] = nw;
    const uint32_t n = e.need()[idx >> 5], nn = tile_needs_visit(e, idx, v) ? (n | bit) : (n & ~bit);
    if (nn != n) e.need()[idx >> 5] = nn;
}
// [ALL] recompute both bitmaps from the map (after bulk writes: clearMap, terrain, state upload)
MPD void rebuild_active(Env &e)
{
    if (e.c.lead()) { e.s->powerQuiet = 0; e.s->powerPure = 0; }
    for (int w = e.c.tid; w < POWER_WORDS; w += e.c.n) {
        uint32_t m = 0, nd = 0;
        for (int b = 0; b < 32; b++) {
            const int v = e.map()[w * 32 + b];
            if (tile_is_active(v)) m |= 1u << b;
            if (tile_needs_visit(e, w * 32 + b, v)) nd |= 1u << b;
        }
        e.act()[w] = m;
        e.need()[w] = nd;
    }
    e.c.sync();
}
// [ALL] the power grid (or newPower) changed: recompute `need` for the active tiles
MPD void refresh_need(Env &e)
{
    for (int w = e.c.tid; w < POWER_WORDS; w += e.c.n) {
        uint32_t m = e.act()[w], nd = 0;
        while (m) {
#if defined(__CUDA_ARCH__)
            const int b = __ffs(m) - 1;
#else
            int b = 0;
            while (!((m >> b) & 1u)) b++;
#endif
            m &= m - 1;
            if (tile_needs_visit(e, w * 32 + b, e.map()[w * 32 + b])) nd |= 1u << b;
        }
        e.need()[w] = nd;
    }
    e.c.sync();
}
// bounds-safe tile store (the reference indexes map[x][y] unchecked in a few places that are
// in-bounds for every reachable state; an out-of-range store is dropped here)
MPD void tset(Env &e, int x, int y, int v) { if (inb(x, y)) mset(e, tix(x, y), v); }
MPD int tget(const Env &e, int x, int y) { return inb(x, y) ? e.map()[tix(x, y)] : 0; }

// ------------------------------------------------------------------ RNG (random.cpp:88-174)
// nextRandom is an unsigned long in the reference but only bits 8..23 are observed and the LCG is
// closed modulo 2^32, so 32 bits of state are output-equivalent.
MPD int sim_random(Env &e)
{
    e.s->rng = e.s->rng * 1103515245u + 12345u;
    return (int)((e.s->rng & 0xffff00u) >> 8);
}
MPD int rnd16(Env &e) { return sim_random(e) & 0xffff; }
MPD int rnd16s(Env &e)
{
    int i = rnd16(e);
    if (i > 0x7fff) i = 0x7fff - i;
    return i;
}
// getRandom(range).  The rejection loop is inline: a call on the serial lane costs far more than the ten instructions
// it saves at each site (out of line: 1.93 M, inline: 2.04 M env-ticks/s; profiles/r02_notes.md); the inline wrapper
// folds the (almost always literal) range into the two numbers the loop needs.
MPD int rnd_reject(Env &e, int maxMultiple)
{
    int r;
    do {
        r = rnd16(e);
    } while (r >= maxMultiple);
    return r;
}
// the modulo stays in the inlined wrapper: with a literal range it folds to a multiply
MPD int16_t rnd(Env &e, int16_t range)
{
    const int range1 = range + 1;
    const int maxMultiple = (0xffff / range1) * range1;
    return (int16_t)(rnd_reject(e, maxMultiple) % range1);
}
MPD int16_t ernd(Env &e, int16_t limit)
{
    int16_t z = rnd(e, limit), x = rnd(e, limit);
    return tmin(z, x);
}
MPD int random_fire(Env &e) { return (T_FIRE + (rnd16(e) & 7)) | ANIMBIT; }
MPD int random_rubble(Env &e) { return (T_RUBBLE + (rnd16(e) & 3)) | BULLBIT; }

// Position::move (position.cpp:140-214) including its two quirks: NORTH_EAST falls through to
// EAST when blocked, SOUTH_WEST always steps then clamps and reports failure.
MPD bool pos_move(int &x, int &y, int dir)
{
    switch (dir) {
    case D_INVALID: return true;
    case D_N: if (y > 0) { y--; return true; } break;
    case D_NE:
        if (x < WORLD_W - 1 && y > 0) { x++; y--; return true; }
        // fall through
    case D_E: if (x < WORLD_W - 1) { x++; return true; } break;
    case D_SE: if (x < WORLD_W - 1 && y < WORLD_H - 1) { x++; y++; return true; } break;
    case D_S: if (y < WORLD_H - 1) { y++; return true; } break;
    case D_SW: x--; y++; break;
    case D_W: if (x > 0) { x--; return true; } break;
    case D_NW: if (x > 0 && y > 0) { x--; y--; return true; } break;
    default: break;
    }
    if (x < 0) x = 0;
    if (x >= WORLD_W) x = WORLD_W - 1;
    if (y < 0) y = 0;
    if (y >= WORLD_H) y = WORLD_H - 1;
    return false;
}
MPD int rot45(int dir, int count) { return ((dir - D_N + count) & 7) + D_N; }

} // namespace mp
