// mp_simt.cu -- the simulation frames with one THREAD per city.
//
// mp_batch.cu gives every city a warp: the RNG-ordered serial walk of the tick runs on lane 0 while 31 lanes
// wait, and the SM fetches one instruction stream per city (ncu r01/r02: 14 of 32 threads active, issue slots
// 27 % busy, the instruction caches the limiter).  Here the same engine headers are compiled with
// Coop::n = 1 (-DMP_SIMT: exactly the single-thread configuration the CPU test harness runs), so the 32 lanes of a
// warp run 32 DIFFERENT cities: lanes whose cities take the same path share its instructions, divergent paths are
// serialised by the hardware.  Results are bit-identical by construction -- a city is still one sequential
// program -- and the state layout in HBM is the one of mp_batch.cu, so a city can be stepped by either kernel from
// one launch to the next (the heavy tail of the cost order stays with the warp-per-city kernels, whose lane-parallel
// flood / power / earthquake paths it needs).
#define MP_SIMT 1
#define mp mp_simt           // the thread-per-env instantiation of the engine lives in its own namespace (no symbol clashes with mp_batch.cu)
#include <cuda_runtime.h>
#include <stdint.h>

#include "mp_env.h"

using namespace mp;

namespace {

constexpr int ENV_STRIDE = (EnvLayout::TOTAL + 127) & ~127;
constexpr int SIMT_THREADS = 128;                       // cities (threads) per CTA
constexpr int SCAL_WORDS = (int)(sizeof(Scalars) / 4);
constexpr int SCAL_PAD_WORDS = SCAL_WORDS;              // 114 words: 8-byte aligned (int64 fields); a 32-bit access of a warp is 2-way conflicted
static_assert(sizeof(Scalars) % 8 == 0, "64-bit fields need an 8-byte stride");
static_assert(sizeof(Scalars) % 4 == 0, "word copies");

struct TArgs {
    uint8_t *blocks; const uint16_t *anim; int n_envs, first, count, animate_before, respect_passes;
    unsigned long long *cycles; unsigned int *cost; const int *perm; int gstride, goff, gcount;
    unsigned int *prof; int prof_off; const uint8_t *mask;
};

__global__ void __launch_bounds__(SIMT_THREADS, 3) k_mp_frames_t(const TArgs a)
{
    const int widx = blockIdx.x * SIMT_THREADS + threadIdx.x;
    const int slot = widx * a.gstride + a.goff;
    if (widx >= a.gcount || slot >= a.n_envs) return;
    const int env = a.perm ? a.perm[slot] : slot;
    if (a.mask && !a.mask[env]) return;
    const long long t_begin = clock64();
    Env e;
    int red[8];
    uint8_t *blk = a.blocks + (size_t)env * ENV_STRIDE;
    bind_env(e, blk);
    e.anim = a.anim;
    e.c.red = red;
    // the city's scalar block in shared memory for the launch
    uint32_t *sw = reinterpret_cast<uint32_t *>(mp_dyn_smem) + (size_t)threadIdx.x * SCAL_PAD_WORDS;
    uint32_t *gs = reinterpret_cast<uint32_t *>(blk + EnvLayout::SCALARS);
    MP_NOUNROLL for (int i = 0; i < SCAL_WORDS; i++) sw[i] = gs[i];
    e.s = reinterpret_cast<Scalars *>(sw);
    if (a.animate_before && e.s->doAnimation && !e.s->tilesAnimated) animate_tiles(e);
    int nrun = a.count;
    if (a.respect_passes) nrun = e.s->simSpeed ? min(a.count, max(0, (int)e.s->simPasses - a.first)) : 0;
    MP_NOUNROLL for (int i = 0; i < nrun; i++) {
        const unsigned int t0 = (unsigned int)clock64();
        sim_loop(e);
        if (a.prof) a.prof[(size_t)env * 256 + a.prof_off + a.first + i] = (unsigned int)clock64() - t0;
    }
    // schedule key: the city's own work (its lanes share the warp's clock): visited tiles, scaled to the order of
    // magnitude of the warp-per-city kernels' key (busy cycles >> 8) so that the same buckets serve both
    if (a.cost) a.cost[env] += e.s->work << 2;
    e.s->work = 0;
    MP_NOUNROLL for (int i = 0; i < SCAL_WORDS; i++) gs[i] = sw[i];
    if (a.cycles) a.cycles[env] += (unsigned long long)(clock64() - t_begin);
}

} // namespace

extern "C" int mps_launch_frames(uint8_t *blocks, const uint16_t *anim, int n_envs, int first, int count, int animate_before,
                                 int respect_passes, unsigned long long *cycles, unsigned int *cost, const int *perm,
                                 int gstride, int goff, int gcount, unsigned int *prof, int prof_off, const uint8_t *mask,
                                 void *stream)
{
    static bool attr_set = false;
    const size_t smem = (size_t)SIMT_THREADS * SCAL_PAD_WORDS * 4;
    if (!attr_set) {
        cudaFuncSetAttribute(k_mp_frames_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    TArgs a;
    a.blocks = blocks; a.anim = anim; a.n_envs = n_envs; a.first = first; a.count = count; a.animate_before = animate_before;
    a.respect_passes = respect_passes; a.cycles = cycles; a.cost = cost; a.perm = perm; a.gstride = gstride; a.goff = goff;
    a.gcount = gcount; a.prof = prof; a.prof_off = prof_off; a.mask = mask;
    k_mp_frames_t<<<(gcount + SIMT_THREADS - 1) / SIMT_THREADS, SIMT_THREADS, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}
