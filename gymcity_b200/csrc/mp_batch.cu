// mp_batch.cu -- batched Micropolis engine on sm_100a: one CTA per city, state resident in HBM as
// one contiguous block per city (mp_state.h EnvLayout), kernels + the mpb_* C-ABI of
// include/gymcity_b200.h.
//
// What this replaces in the reference: the C++ `Micropolis` engine object as the gym path drives
// it through SWIG (gym_city/envs/corecontrol.py:36-357) -- construction, clearMap/generateMap,
// toolDown, setFunds, simTick / tickEngine, and the per-cell getters used to read state back.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gymcity_b200.h"
#include "mp_env.h"
#include "mp_snapshot_io.h"

using namespace mp;

namespace {

constexpr int MP_THREADS = 128;
constexpr int ENV_STRIDE = (EnvLayout::TOTAL + 127) & ~127;

struct MpBatch {
    int n_envs, device;
    uint8_t *blocks;          // [n_envs * ENV_STRIDE]
    uint16_t *anim;           // [1024]
    int32_t *i32_a, *i32_b;   // staging [n_envs * 4]
    uint8_t *mask;            // [n_envs]
    uint8_t *host_blk;        // pinned, one env block
    int32_t *host_i32;        // pinned staging [n_envs * 4]
    char err[256];
};

__device__ __forceinline__ void bind_device_env(Env &e, uint8_t *blocks, const uint16_t *anim, int *red)
{
    bind_env(e, blocks + (size_t)blockIdx.x * ENV_STRIDE);
    e.anim = anim;
    e.c.tid = threadIdx.x;
    e.c.n = blockDim.x;
    e.c.red = red;
}

__global__ void __launch_bounds__(MP_THREADS) k_mp_construct(uint8_t *blocks, const uint16_t *anim)
{
    __shared__ int red[8];
    Env e;
    bind_device_env(e, blocks, anim, red);
    env_construct(e, blocks + (size_t)blockIdx.x * ENV_STRIDE);
}

// mode 0: clearMap; mode 1: generateSomeCity(terrain_seed) then seedRandom(sim_seed)
__global__ void __launch_bounds__(MP_THREADS)
k_mp_reset(uint8_t *blocks, const uint16_t *anim, const int32_t *terrain_seeds,
           const int32_t *sim_seeds, const uint8_t *mask, int mode)
{
    __shared__ int red[8];
    if (mask && !mask[blockIdx.x]) return;
    Env e;
    bind_device_env(e, blocks, anim, red);
    if (mode == 0) clear_map(e);
    else generate_some_city(e, terrain_seeds[blockIdx.x], sim_seeds[blockIdx.x]);
}

__global__ void __launch_bounds__(MP_THREADS)
k_mp_tool(uint8_t *blocks, const uint16_t *anim, const int32_t *txy, int32_t *results)
{
    __shared__ int red[8];
    Env e;
    bind_device_env(e, blocks, anim, red);
    if (threadIdx.x == 0) {
        const int32_t *a = txy + (size_t)blockIdx.x * 3;
        int r = (a[0] < 0) ? (int)TR_FAILED : tool_down(e, a[0], a[1], a[2]);
        if (results) results[blockIdx.x] = r;
    }
}

// mode 0: n x simLoop; mode 1: MicropolisControl.simTick (2 x simPasses frames + animateTiles)
__global__ void __launch_bounds__(MP_THREADS)
k_mp_tick(uint8_t *blocks, const uint16_t *anim, int mode, int n)
{
    __shared__ int red[8];
    Env e;
    bind_device_env(e, blocks, anim, red);
    if (mode == 0) {
        for (int i = 0; i < n; i++) sim_loop(e);
    } else if (mode == 1) {
        control_sim_tick(e);
    } else {
        sim_tick(e);
        if (mode == 3 && e.s->doAnimation && !e.s->tilesAnimated) animate_tiles(e);
    }
}

__global__ void __launch_bounds__(MP_THREADS)
k_mp_call(uint8_t *blocks, const uint16_t *anim, int fn, int a, int b)
{
    __shared__ int red[8];
    Env e;
    bind_device_env(e, blocks, anim, red);
    env_call(e, fn, a, b);
}

__global__ void k_mp_set_scalar(uint8_t *blocks, int n, int which, const int32_t *vals, int64_t v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Scalars *s = (Scalars *)(blocks + (size_t)i * ENV_STRIDE + EnvLayout::SCALARS);
    switch (which) {
    case 0: s->rng = (uint32_t)vals[i]; break;
    case 1: s->totalFunds = (int)v; break;
    case 2: s->simPasses = (int)v; break;
    case 3: s->enableDisasters = (uint8_t)v; break;
    default: break;
    }
}

int fail(MpBatch *b, const char *what, cudaError_t e)
{
    snprintf(b->err, sizeof(b->err), "%s: %s", what, cudaGetErrorString(e));
    return -1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(b, #call, e_); } while (0)

int fetch_env(MpBatch *b, int env, Env &e)
{
    if (env < 0 || env >= b->n_envs) { snprintf(b->err, sizeof(b->err), "env %d out of range", env); return -1; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(b->host_blk, b->blocks + (size_t)env * ENV_STRIDE, EnvLayout::TOTAL, cudaMemcpyDeviceToHost));
    bind_env(e, b->host_blk);
    return 0;
}
int store_env(MpBatch *b, int env)
{
    CK(cudaMemcpy(b->blocks + (size_t)env * ENV_STRIDE, b->host_blk, EnvLayout::TOTAL, cudaMemcpyHostToDevice));
    return 0;
}

} // namespace

extern "C" {

mpb_handle mpb_create(int n_envs, int device)
{
    if (n_envs <= 0) return NULL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return NULL;
    cudaSetDevice(device);
    MpBatch *b = (MpBatch *)calloc(1, sizeof(MpBatch));
    b->n_envs = n_envs; b->device = device;
    uint16_t anim[1024];
    for (int i = 0; i < 1024; i++) anim[i] = (uint16_t)i;
    for (int i = 0; i < kAnimPairCount; i++) anim[kAnimPairs[i][0]] = kAnimPairs[i][1];
    bool ok = cudaMalloc(&b->blocks, (size_t)n_envs * ENV_STRIDE) == cudaSuccess
        && cudaMalloc(&b->anim, sizeof(anim)) == cudaSuccess
        && cudaMalloc(&b->i32_a, (size_t)n_envs * 16) == cudaSuccess
        && cudaMalloc(&b->i32_b, (size_t)n_envs * 16) == cudaSuccess
        && cudaMalloc(&b->mask, n_envs) == cudaSuccess
        && cudaMallocHost(&b->host_blk, ENV_STRIDE) == cudaSuccess
        && cudaMallocHost(&b->host_i32, (size_t)n_envs * 16) == cudaSuccess;
    if (!ok) { mpb_destroy(b); return NULL; }
    cudaMemcpy(b->anim, anim, sizeof(anim), cudaMemcpyHostToDevice);
    k_mp_construct<<<n_envs, MP_THREADS>>>(b->blocks, b->anim);
    if (cudaDeviceSynchronize() != cudaSuccess) { mpb_destroy(b); return NULL; }
    return b;
}

void mpb_destroy(mpb_handle h)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return;
    cudaSetDevice(b->device);
    cudaFree(b->blocks); cudaFree(b->anim); cudaFree(b->i32_a); cudaFree(b->i32_b); cudaFree(b->mask);
    if (b->host_blk) cudaFreeHost(b->host_blk);
    if (b->host_i32) cudaFreeHost(b->host_i32);
    free(b);
}

const char *mpb_last_error(mpb_handle h) { return h ? ((MpBatch *)h)->err : "null handle"; }
int mpb_num_envs(mpb_handle h) { return h ? ((MpBatch *)h)->n_envs : -1; }
int mpb_env_stride(void) { return ENV_STRIDE; }
int mpb_state_bytes(void) { return EnvLayout::PERSISTENT_END; }
void *mpb_blocks(mpb_handle h) { return h ? ((MpBatch *)h)->blocks : NULL; }

int mpb_seed(mpb_handle h, const int32_t *seeds_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !seeds_host) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->host_i32, seeds_host, (size_t)b->n_envs * 4);
    CK(cudaMemcpyAsync(b->i32_a, b->host_i32, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
    k_mp_set_scalar<<<(b->n_envs + 127) / 128, 128, 0, st>>>(b->blocks, b->n_envs, 0, b->i32_a, 0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

static int set_all(MpBatch *b, int which, int64_t v, cudaStream_t st)
{
    k_mp_set_scalar<<<(b->n_envs + 127) / 128, 128, 0, st>>>(b->blocks, b->n_envs, which, NULL, v);
    CK(cudaGetLastError());
    return 0;
}
int mpb_set_funds(mpb_handle h, int funds, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    return set_all(b, 1, funds, (cudaStream_t)stream);
}
int mpb_set_passes(mpb_handle h, int passes, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    return set_all(b, 2, passes, (cudaStream_t)stream);
}
int mpb_set_disasters(mpb_handle h, int enabled, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    return set_all(b, 3, enabled ? 1 : 0, (cudaStream_t)stream);
}

static int upload_mask(MpBatch *b, const uint8_t *mask_host, cudaStream_t st)
{
    if (!mask_host) return 0;
    CK(cudaMemcpyAsync(b->mask, mask_host, b->n_envs, cudaMemcpyHostToDevice, st));
    return 0;
}

int mpb_clear_map(mpb_handle h, const uint8_t *mask_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (upload_mask(b, mask_host, st)) return -1;
    k_mp_reset<<<b->n_envs, MP_THREADS, 0, st>>>(b->blocks, b->anim, NULL, NULL, mask_host ? b->mask : NULL, 0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_generate(mpb_handle h, const int32_t *terrain_seeds_host, const int32_t *sim_seeds_host,
                 const uint8_t *mask_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !terrain_seeds_host || !sim_seeds_host) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (upload_mask(b, mask_host, st)) return -1;
    memcpy(b->host_i32, terrain_seeds_host, (size_t)b->n_envs * 4);
    memcpy(b->host_i32 + b->n_envs, sim_seeds_host, (size_t)b->n_envs * 4);
    CK(cudaMemcpyAsync(b->i32_a, b->host_i32, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b->i32_b, b->host_i32 + b->n_envs, (size_t)b->n_envs * 4, cudaMemcpyHostToDevice, st));
    k_mp_reset<<<b->n_envs, MP_THREADS, 0, st>>>(b->blocks, b->anim, b->i32_a, b->i32_b,
                                                mask_host ? b->mask : NULL, 1);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

int mpb_tool(mpb_handle h, const int32_t *tool_xy_dev, int32_t *results_dev, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !tool_xy_dev) return -1;
    cudaSetDevice(b->device);
    k_mp_tool<<<b->n_envs, MP_THREADS, 0, (cudaStream_t)stream>>>(b->blocks, b->anim, tool_xy_dev, results_dev);
    CK(cudaGetLastError());
    return 0;
}

int mpb_tool_host(mpb_handle h, const int32_t *tool_xy_host, int32_t *results_host, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !tool_xy_host) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->host_i32, tool_xy_host, (size_t)b->n_envs * 12);
    CK(cudaMemcpyAsync(b->i32_a, b->host_i32, (size_t)b->n_envs * 12, cudaMemcpyHostToDevice, st));
    if (mpb_tool(h, b->i32_a, b->i32_b, stream)) return -1;
    CK(cudaMemcpyAsync(b->host_i32, b->i32_b, (size_t)b->n_envs * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (results_host) memcpy(results_host, b->host_i32, (size_t)b->n_envs * 4);
    return 0;
}

int mpb_sim_frames(mpb_handle h, int n, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || n < 0) return -1;
    cudaSetDevice(b->device);
    k_mp_tick<<<b->n_envs, MP_THREADS, 0, (cudaStream_t)stream>>>(b->blocks, b->anim, 0, n);
    CK(cudaGetLastError());
    return 0;
}

int mpb_control_sim_tick(mpb_handle h, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    k_mp_tick<<<b->n_envs, MP_THREADS, 0, (cudaStream_t)stream>>>(b->blocks, b->anim, 1, 0);
    CK(cudaGetLastError());
    return 0;
}

int mpb_sim_tick(mpb_handle h, int animate, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    k_mp_tick<<<b->n_envs, MP_THREADS, 0, (cudaStream_t)stream>>>(b->blocks, b->anim, animate ? 3 : 2, 0);
    CK(cudaGetLastError());
    return 0;
}

int mpb_call(mpb_handle h, int fn, int a, int bb, void *stream)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    k_mp_call<<<b->n_envs, MP_THREADS, 0, (cudaStream_t)stream>>>(b->blocks, b->anim, fn, a, bb);
    CK(cudaGetLastError());
    return 0;
}

int mpb_sync(mpb_handle h)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    CK(cudaDeviceSynchronize());
    return 0;
}

int mpb_get_field(mpb_handle h, int env, int field, void *out_host, int cap)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    return snapshot_get_field(e, field, out_host, cap);
}

int mpb_set_field(mpb_handle h, int env, int field, const void *in_host, int nbytes)
{
    MpBatch *b = (MpBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    if (snapshot_set_field(e, field, in_host, nbytes)) return -1;
    return store_env(b, env);
}

int mpb_get_scalars(mpb_handle h, int env, mp_scalars *out)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !out) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    snapshot_get_scalars(e, out);
    return 0;
}

int mpb_set_scalars(mpb_handle h, int env, const mp_scalars *in)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !in) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    snapshot_set_scalars(e, in);
    return store_env(b, env);
}

int mpb_get_sprites(mpb_handle h, int env, mp_sprite *out, int cap)
{
    MpBatch *b = (MpBatch *)h;
    if (!b || !out) return -1;
    cudaSetDevice(b->device);
    Env e;
    if (fetch_env(b, env, e)) return -1;
    return snapshot_get_sprites(e, out, cap);
}

} // extern "C"
