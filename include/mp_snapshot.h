/* mp_snapshot.h -- flat snapshot of one Micropolis engine's scalar state.
 *
 * Shared by the CUDA library (mpb_get_scalars) and the oracle shim
 * (mpref_get_scalars) so parity tests can diff both sides field by field.
 * Every field is widened to int64 / float regardless of the engine's internal type; the
 * engines themselves keep the reference's exact widths (short / int / long), see
 * reference micropolis.h:1046-1230, 1602-1700, 2153-2330.
 *
 * The field list is an X-macro so that the C struct, the name table and the Python
 * ctypes mirror can never drift apart.
 */
#ifndef MP_SNAPSHOT_H
#define MP_SNAPSHOT_H

#include <stdint.h>

#define MP_SCALAR_INT_FIELDS(X) \
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
    X(powerStackPointer) X(nextRandom) \
    X(cityCenterX) X(cityCenterY) X(pollutionMaxX) X(pollutionMaxY) \
    X(crimeMaxX) X(crimeMaxY) \
    X(crimeRamp) X(pollutionRamp) X(resCap) X(comCap) X(indCap) X(cashFlow) \
    X(disasterEvent) X(disasterWait) X(scoreType) X(scoreWait) \
    X(poweredZoneCount) X(unpoweredZoneCount) X(newPower) X(cityTaxAverage) \
    X(simCycle) X(phaseCycle) X(speedCycle) X(doInitialEval) \
    X(resValve) X(comValve) X(indValve) \
    X(spriteCycle) X(totalFunds) X(autoBulldoze) X(autoBudget) X(gameLevel) \
    X(initSimLoad) X(scenario) X(simSpeed) X(simPasses) X(enableDisasters) \
    X(doAnimation) X(tilesAnimated) X(trafMaxX) X(trafMaxY) X(absDist) \
    X(numSprites)

#define MP_SCALAR_FLOAT_FIELDS(X) \
    X(roadPercent) X(policePercent) X(firePercent) X(externalMarket)

#define MP_PROBNUM 10
#define MP_PROBLEM_COMPLAINTS 4

typedef struct mp_scalars {
#define X(n) int64_t n;
    MP_SCALAR_INT_FIELDS(X)
#undef X
    int64_t problemVotes[MP_PROBNUM];
    int64_t problemOrder[MP_PROBLEM_COMPLAINTS];
#define X(n) float n;
    MP_SCALAR_FLOAT_FIELDS(X)
#undef X
} mp_scalars;

/* One sprite, in the reference's list order (head first); reference micropolis.h:981-1018.
 * `name` and pointers are dropped; everything else is copied verbatim. */
typedef struct mp_sprite {
    int32_t type, frame, x, y, width, height, xOffset, yOffset, xHot, yHot;
    int32_t origX, origY, destX, destY, count, soundCount, dir, newDir;
    int32_t step, flag, control, turn, accel, speed;
} mp_sprite;

#define MP_MAX_SPRITES 64

/* Field ids for the array getters/setters (mpb_get_field / mpref_get_field). */
enum mp_field {
    MPF_MAP = 0,            /* uint16[120*100], column-major x*100+y          */
    MPF_POP_DENSITY,        /* uint8 [60*50]                                  */
    MPF_TRAFFIC_DENSITY,    /* uint8 [60*50]                                  */
    MPF_POLLUTION_DENSITY,  /* uint8 [60*50]                                  */
    MPF_LAND_VALUE,         /* uint8 [60*50]                                  */
    MPF_CRIME_RATE,         /* uint8 [60*50]                                  */
    MPF_TERRAIN_DENSITY,    /* uint8 [30*25]                                  */
    MPF_POWER_GRID,         /* uint8 [120*100] (0/1, unpacked)                */
    MPF_RATE_OF_GROWTH,     /* int16 [15*13]                                  */
    MPF_FIRE_STATION,       /* int16 [15*13]                                  */
    MPF_FIRE_STATION_EFFECT,/* int16 [15*13]                                  */
    MPF_POLICE_STATION,     /* int16 [15*13]                                  */
    MPF_POLICE_STATION_EFFECT, /* int16 [15*13]                               */
    MPF_COM_RATE,           /* int16 [15*13]                                  */
    MPF_RES_HIST,           /* int16 [240]                                    */
    MPF_COM_HIST,           /* int16 [240]                                    */
    MPF_IND_HIST,           /* int16 [240]                                    */
    MPF_MONEY_HIST,         /* int16 [240]                                    */
    MPF_POLLUTION_HIST,     /* int16 [240]                                    */
    MPF_CRIME_HIST,         /* int16 [240]                                    */
    MPF_MISC_HIST,          /* int16 [120]                                    */
    MPF_POWER_STACK,        /* int16 [3000*2] (x,y pairs; entries 1..pointer) */
    MPF_COUNT
};

#endif /* MP_SNAPSHOT_H */
