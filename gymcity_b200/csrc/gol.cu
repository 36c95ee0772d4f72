// gol.cu -- batched 1-player Game of Life step for sm_100a, behind the C-ABI in
// include/gymcity_b200.h (golb_*).
//
// What it replaces (reference game_of_life/envs/):
//   multi_env.py:170-232  GoLMultiEnv.step   : XOR-toggle one cell per env, tick, pop, reward, obs
//   world_pytorch.py:91-135 World._tick/GoLu : torus B3/S23 (conv value v = n + 9*self, alive iff
//                                              v in {3, 11, 12})
//   world.py:29-56,108-130 + env.py:125-182   : single-env object world incl. its stale-neighbour
//                                              behaviour (see golb_classic_step below)
//
// Layout in HBM: one uint32 per board row (bit y of row x = cell [x][y]), rows of one board
// contiguous, boards contiguous: rows[env * W + x].  A 16x16 board is 64 B.  One thread owns one
// row; the rows above/below come from warp shuffles when W divides 32 (the benchmark sizes
// 8/16/32), from shared memory otherwise.  Neighbour counts are computed for all W cells of a row
// at once with a bit-sliced adder (no per-cell loop), population with __popc.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../include/gymcity_b200.h"

namespace {

struct GolBatch {
    int n_envs, width, max_step, device;
    int classic;                 // 1: single-env semantics (env.py / world.py), 0: multi_env.py
    uint32_t *rows;              // [n_envs * W]
    uint32_t *frozen;            // classic only: [n_envs * W * 8] per-direction "reference is stale"
    uint32_t *frozen_val;        // classic only: [n_envs * W * 8] value seen through a stale reference
    float *obs;                  // multi: [n, 2, W, W] fp32; classic: unused
    uint8_t *obs_u8;             // classic: [n, 1, W, W] uint8
    float *reward;               // [n]
    int32_t *pop;                // [n]
    uint8_t *done;               // [n]
    int64_t *act_dev;            // staging for the host-buffer entry point
    float trg_pop, max_loss, scalar_plane;
    int step_count;
    char err[256];
    // pinned host staging for golb_step_host
    int64_t *h_act;
    float *h_reward;
    cudaEvent_t ev;
};

__device__ __forceinline__ uint32_t rotl_w(uint32_t r, int w, uint32_t mask)
{
    return ((r << 1) | (r >> (w - 1))) & mask;
}
__device__ __forceinline__ uint32_t rotr_w(uint32_t r, int w, uint32_t mask)
{
    return ((r >> 1) | (r << (w - 1))) & mask;
}

// Next-generation row from the row itself and its torus neighbours above/below.
// Bit-sliced: s0/s1 = 2-bit horizontal sums of each of the three rows, then a carry-save add of
// the eight neighbours into count bits b0..b3; alive' = (n == 3) | (self & n == 2), which is the
// rule table v in {3, 11, 12} of world_pytorch.py:110-135 with v = n + 9*self.
__device__ __forceinline__ uint32_t life_row(uint32_t up, uint32_t me, uint32_t dn, int w,
                                             uint32_t mask)
{
    uint32_t ul = rotl_w(up, w, mask), ur = rotr_w(up, w, mask);
    uint32_t ml = rotl_w(me, w, mask), mr = rotr_w(me, w, mask);
    uint32_t dl = rotl_w(dn, w, mask), dr = rotr_w(dn, w, mask);
    // three-input adders: up triple, down triple; two-input: middle pair
    uint32_t u0 = ul ^ up ^ ur, u1 = (ul & up) | (ur & (ul ^ up));
    uint32_t d0 = dl ^ dn ^ dr, d1 = (dl & dn) | (dr & (dl ^ dn));
    uint32_t m0 = ml ^ mr, m1 = ml & mr;
    // add the three 2-bit numbers
    uint32_t a0 = u0 ^ d0 ^ m0;
    uint32_t c0 = (u0 & d0) | (m0 & (u0 ^ d0));          // carry into bit 1
    uint32_t t1 = u1 ^ d1 ^ m1;
    uint32_t c1 = (u1 & d1) | (m1 & (u1 ^ d1));          // carry into bit 2 from the bit-1 column
    uint32_t a1 = t1 ^ c0;
    uint32_t c1b = t1 & c0;                               // second carry into bit 2
    uint32_t a2 = c1 ^ c1b;
    uint32_t a3 = c1 & c1b;
    uint32_t two_or_three = a1 & ~a2 & ~a3;               // n in {2, 3}
    return two_or_three & (a0 | me) & mask;
}

// One thread per row.  W | 32 path: shuffle halo exchange inside the W-lane group.
// T ticks per launch (golb_step_n): the rows stay in registers from tick to tick, tick t takes its actions from
// actions[t * n_envs + env]; per-tick observations / rewards go to obs_seq[t] / reward_seq[t] when given, the
// last tick's always to the handle's own buffers -- the state afterwards is that of T golb_step calls.  At
// BASELINE's 4 096 boards one tick is ~1 us of GPU work: a launch per tick is launch-bound (14 us), T per launch
// is not.
template <bool SHFL>
__global__ void __launch_bounds__(256)
k_gol_step(uint32_t *__restrict__ rows, const int64_t *__restrict__ actions,
           float *__restrict__ obs, float *__restrict__ reward, int32_t *__restrict__ pop_out,
           uint8_t *__restrict__ done, int n_envs, int w, float trg_pop, float max_loss,
           float denom, int step_count, int max_step, int T, float *__restrict__ obs_seq,
           float *__restrict__ reward_seq, float scalar_plane)
{
    extern __shared__ uint32_t s_rows[];
    const uint32_t mask = (w == 32) ? 0xffffffffu : ((1u << w) - 1u);
    const int envs_per_cta = blockDim.x / w;
    const int le = threadIdx.x / w;                 // env within the CTA
    const int x = threadIdx.x - le * w;             // row within the board
    const int env = blockIdx.x * envs_per_cta + le;
    const bool live = (le < envs_per_cta) && (env < n_envs);
    const size_t plane = (size_t)w * w;

    uint32_t me = live ? rows[(size_t)env * w + x] : 0u;
    for (int t = 0; t < T; t++) {
        if (live) {
            // multi_env.py:179,243-253: new_state = state ^ one_hot(a), a = x * W + y
            int64_t a = actions[(size_t)t * n_envs + env];
            int ax = (int)(a / w), ay = (int)(a - (int64_t)ax * w);
            if (ax == x) me ^= (1u << ay);
        }
        uint32_t up, dn;
        if (SHFL) {
            // rows of one board sit in one aligned group of w lanes
            const int lane = threadIdx.x & 31;
            const int base = lane - x;
            up = __shfl_sync(0xffffffffu, me, base + ((x + w - 1) & (w - 1)));
            dn = __shfl_sync(0xffffffffu, me, base + ((x + 1) & (w - 1)));
        } else {
            __syncthreads();
            s_rows[threadIdx.x] = me;
            __syncthreads();
            int xu = (x == 0) ? w - 1 : x - 1, xd = (x == w - 1) ? 0 : x + 1;
            up = live ? s_rows[le * w + xu] : 0;
            dn = live ? s_rows[le * w + xd] : 0;
        }
        const uint32_t nxt = life_row(up, me, dn, w, mask);
        int p = __popc(nxt);
        if (SHFL) {
            for (int o = w >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        } else {
            __syncthreads();
            s_rows[threadIdx.x] = (uint32_t)p;
            __syncthreads();
            if (live && x == 0) {
                int s = 0;
                for (int i = 0; i < w; i++) s += (int)s_rows[le * w + i];
                p = s;
            }
        }
        const bool last = t == T - 1;
        // where this tick's observation goes: the caller's sequence buffer, and the handle's own for the last tick
        for (int dst = 0; dst < 2; dst++) {
            float *ob = dst == 0 ? (obs_seq ? obs_seq + (size_t)t * n_envs * 2 * plane : (float *)0)
                                 : (last ? obs : (float *)0);
            if (!ob) continue;
            const bool with_plane0 = dst == 0;      // the handle's plane 0 is written at set_params; a sequence buffer gets it here
            if (SHFL && w == 16) {
                // 16x16 boards: a warp holds two boards.  Instead of every lane writing its own row (four 16-byte
                // stores 64 bytes apart from their neighbours'), the rows are redistributed by shuffle so that
                // each store instruction writes 16 consecutive float4 per board: chunk c = k * 16 + j of a board's
                // 1 KB state plane is quarter (j & 3) of row (k * 4 + j / 4).
                const int lane = threadIdx.x & 31, j = lane & 15, half = lane & 16;
                float4 *o = reinterpret_cast<float4 *>(ob + ((size_t)env * 2 + 1) * 256);
                const float4 sp = make_float4(scalar_plane, scalar_plane, scalar_plane, scalar_plane);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t r = __shfl_sync(0xffffffffu, nxt, half + k * 4 + (j >> 2)) >> ((j & 3) * 4);
                    float4 v;
                    v.x = (float)(r & 1u);
                    v.y = (float)((r >> 1) & 1u);
                    v.z = (float)((r >> 2) & 1u);
                    v.w = (float)((r >> 3) & 1u);
                    if (live) {
                        __stcs(&o[k * 16 + j], v);
                        if (with_plane0) __stcs(&o[k * 16 + j - 64], sp);
                    }
                }
            } else if (live) {
                // multi_env.py:235-240: obs = cat(scalar plane, state)
                float4 *o = reinterpret_cast<float4 *>(ob + ((size_t)env * 2 + 1) * plane + (size_t)x * w);
                if ((w & 3) == 0) {
                    for (int y = 0; y < w; y += 4) {
                        float4 v;
                        v.x = (float)((nxt >> y) & 1u);
                        v.y = (float)((nxt >> (y + 1)) & 1u);
                        v.z = (float)((nxt >> (y + 2)) & 1u);
                        v.w = (float)((nxt >> (y + 3)) & 1u);
                        o[y >> 2] = v;
                    }
                } else {
                    float *of = reinterpret_cast<float *>(o);
                    for (int y = 0; y < w; y++) of[y] = (float)((nxt >> y) & 1u);
                }
                if (with_plane0) {
                    float *o0 = ob + (size_t)env * 2 * plane + (size_t)x * w;
                    for (int y = 0; y < w; y++) o0[y] = scalar_plane;
                }
            }
        }
        if (live && x == 0) {
            // multi_env.py:203-212, fp32 like torch: (max_loss - |trg - pop|) * 100 / (max_loss * max_step * num_proc)
            float loss = fabsf(__fsub_rn(trg_pop, (float)p));
            float r = __fdiv_rn(__fmul_rn(__fsub_rn(max_loss, loss), 100.0f), denom);
            if (reward_seq) reward_seq[(size_t)t * n_envs + env] = r;
            if (last) {
                reward[env] = r;
                pop_out[env] = p;
                // multi_env.py:213-216: terminal is all-true when step_count == max_step (checked before ++)
                done[env] = (uint8_t)(step_count + t == max_step);
            }
        }
        me = nxt;
    }
    if (live) rows[(size_t)env * w + x] = me;
}

// Classic single-env world (world.py).  build_cell() replaces the Cell object at (x,y) while the
// eight neighbours keep their cached reference to the old object (world.py:108-112,119-130), whose
// `alive` never changes again.  So each cell sees each neighbour either live or through a frozen
// value.  frozen[d] / frozen_val[d] are bit-rows per direction d (0..7) in the order of
// world.py:16-20 cached_directions.  One thread per env (W small, the env is the unit here).
__global__ void k_gol_classic_step(uint32_t *__restrict__ rows, uint32_t *__restrict__ frozen,
                                   uint32_t *__restrict__ fval, const int64_t *__restrict__ actions,
                                   uint8_t *__restrict__ obs, float *__restrict__ reward,
                                   int32_t *__restrict__ pop_out, uint8_t *__restrict__ done,
                                   int n_envs, int w, int max_step, int step_count)
{
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n_envs) return;
    const uint32_t mask = (w == 32) ? 0xffffffffu : ((1u << w) - 1u);
    uint32_t *r = rows + (size_t)env * w;
    uint32_t *fz = frozen + (size_t)env * w * 8;
    uint32_t *fv = fval + (size_t)env * w * 8;
    const int dx[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    const int dy[8] = {1, 1, 1, 0, 0, -1, -1, -1};

    // env.py:141-152: reward = state.sum() BEFORE the build; a < 0 deletes (player only)
    int64_t a = actions[env];
    bool del = a < 0;
    if (del) a = -a;
    // an action outside the board (intsToActions raises IndexError in the reference; host-side inputs are
    // rejected by the Python wrapper) builds nothing here: the tick still runs
    const bool in_range = a < (int64_t)w * w;
    int ax = in_range ? (int)(a / w) : 0, ay = in_range ? (int)(a - (int64_t)ax * w) : 0;
    int pre = 0;
    for (int x = 0; x < w; x++) pre += __popc(r[x]);
    // build_cell: neighbours q of p=(ax,ay) whose reference to p is still live now freeze the old value
    uint32_t old = (r[ax] >> ay) & 1u;
    // env.py:147-152: a player delete happens BEFORE reward = state.sum(); a build after it
    if (del && in_range && old) pre--;
    for (int d = 0; d < 8 && in_range; d++) {
        // q sees p in direction d  <=>  q + (dx,dy) = p
        int qx = (ax - dx[d] + w) % w, qy = (ay - dy[d] + w) % w;
        uint32_t bit = 1u << qy;
        if (!(fz[qx * 8 + d] & bit)) {
            fz[qx * 8 + d] |= bit;
            fv[qx * 8 + d] = (fv[qx * 8 + d] & ~bit) | (old ? bit : 0u);
        }
        // the new Cell at p has neighbours=None: all eight references are rebuilt live
        fz[ax * 8 + d] &= ~(1u << ay);
    }
    if (in_range) r[ax] = del ? (r[ax] & ~(1u << ay)) : (r[ax] | (1u << ay));

    // world.py:29-56 _tick with per-direction visibility
    uint32_t nxt_rows[32];
    for (int x = 0; x < w; x++) {
        uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        for (int d = 0; d < 8; d++) {
            uint32_t src = r[(x + dx[d] + w) % w];
            // neighbour at y+dy: bit y of `seen` = src bit (y+dy) mod w
            uint32_t seen = (dy[d] == 1) ? rotr_w(src, w, mask)
                          : (dy[d] == -1) ? rotl_w(src, w, mask) : src;
            uint32_t f = fz[x * 8 + d];
            seen = (seen & ~f) | (fv[x * 8 + d] & f);
            uint32_t c = b0 & seen; b0 ^= seen;
            uint32_t c2 = b1 & c;   b1 ^= c;
            uint32_t c3 = b2 & c2;  b2 ^= c2;
            b3 ^= c3;
        }
        uint32_t me = r[x];
        uint32_t n3 = b0 & b1 & ~b2 & ~b3, n2 = ~b0 & b1 & ~b2 & ~b3;
        nxt_rows[x] = (n3 | (me & n2)) & mask;
    }
    int p = 0;
    for (int x = 0; x < w; x++) {
        r[x] = nxt_rows[x];
        p += __popc(nxt_rows[x]);
        for (int y = 0; y < w; y++) obs[((size_t)env * w + x) * w + y] = (nxt_rows[x] >> y) & 1u;
    }
    // env.py:171-174: terminal = step_count == max_step or raw reward < 2
    done[env] = (uint8_t)((step_count == max_step) || (pre < 2));
    reward[env] = (float)((double)pre * 100.0 / ((double)max_step * w * w));
    pop_out[env] = p;
}

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// world_pytorch.py:75-82 repopulate_cells: alive w.p. prob_life; counter-based hash instead of
// torch's generator (the reference's own stream is not reproducible across devices either).
__global__ void k_gol_populate(uint32_t *rows, int n_rows, int w, uint32_t seed, uint32_t thresh24)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    uint32_t r = 0;
    for (int y = 0; y < w; y++) {
        uint32_t h = mix32(mix32(seed ^ 0x9e3779b9u) + (uint32_t)i * 0x85ebca6bu + (uint32_t)y * 0xc2b2ae35u);
        if ((h >> 8) < thresh24) r |= 1u << y;
    }
    rows[i] = r;
}

__global__ void k_gol_write_obs(const uint32_t *__restrict__ rows, float *obs, uint8_t *obs_u8,
                                int n_envs, int w, float scalar_plane)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per cell
    int cells = w * w;
    if (i >= n_envs * cells) return;
    int env = i / cells, c = i - env * cells, x = c / w, y = c - x * w;
    uint32_t b = (rows[(size_t)env * w + x] >> y) & 1u;
    if (obs) {
        obs[((size_t)env * 2) * cells + c] = scalar_plane;
        obs[((size_t)env * 2 + 1) * cells + c] = (float)b;
    }
    if (obs_u8) obs_u8[(size_t)env * cells + c] = (uint8_t)b;
}

__global__ void k_gol_pack(const uint8_t *__restrict__ cells, uint32_t *rows, int n_rows, int w)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    uint32_t r = 0;
    for (int y = 0; y < w; y++) if (cells[(size_t)i * w + y]) r |= 1u << y;
    rows[i] = r;
}

int fail(GolBatch *b, const char *what, cudaError_t e)
{
    snprintf(b->err, sizeof(b->err), "%s: %s", what, cudaGetErrorString(e));
    return -1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(b, #call, e_); } while (0)

int refresh_obs(GolBatch *b, cudaStream_t st)
{
    int cells = b->n_envs * b->width * b->width;
    k_gol_write_obs<<<(cells + 255) / 256, 256, 0, st>>>(b->rows, b->obs, b->obs_u8, b->n_envs,
                                                         b->width, b->scalar_plane);
    CK(cudaGetLastError());
    return 0;
}

} // namespace

extern "C" {

golb_handle golb_create(int n_envs, int width, int max_step, int classic, int device)
{
    if (n_envs <= 0 || width < 3 || width > 32 || max_step <= 0) return NULL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return NULL;
    GolBatch *b = (GolBatch *)calloc(1, sizeof(GolBatch));
    b->n_envs = n_envs; b->width = width; b->max_step = max_step; b->device = device;
    b->classic = classic;
    cudaSetDevice(device);
    size_t nrows = (size_t)n_envs * width, cells = nrows * width;
    bool ok = cudaMalloc(&b->rows, nrows * 4) == cudaSuccess
        && cudaMalloc(&b->reward, (size_t)n_envs * 4) == cudaSuccess
        && cudaMalloc(&b->pop, (size_t)n_envs * 4) == cudaSuccess
        && cudaMalloc(&b->done, n_envs) == cudaSuccess
        && cudaMalloc(&b->act_dev, (size_t)n_envs * 8) == cudaSuccess
        && cudaMallocHost(&b->h_act, (size_t)n_envs * 8) == cudaSuccess
        && cudaMallocHost(&b->h_reward, (size_t)n_envs * 4) == cudaSuccess
        && cudaEventCreateWithFlags(&b->ev, cudaEventDisableTiming) == cudaSuccess;
    if (ok && classic) {
        ok = cudaMalloc(&b->frozen, nrows * 8 * 4) == cudaSuccess
          && cudaMalloc(&b->frozen_val, nrows * 8 * 4) == cudaSuccess
          && cudaMalloc(&b->obs_u8, cells) == cudaSuccess;
        if (ok) { cudaMemset(b->frozen, 0, nrows * 32); cudaMemset(b->frozen_val, 0, nrows * 32); }
    } else if (ok) {
        ok = cudaMalloc(&b->obs, cells * 2 * 4) == cudaSuccess;
    }
    if (!ok) { golb_destroy(b); return NULL; }
    cudaMemset(b->rows, 0, nrows * 4);
    cudaMemset(b->done, 0, n_envs);
    // multi_env.py:57-72: max_pop = W*W*2/3 is both the target and the loss range
    float max_pop = (float)((double)width * width * 2 / 3);
    b->trg_pop = max_pop; b->max_loss = max_pop; b->scalar_plane = max_pop / max_pop;
    refresh_obs(b, 0);
    cudaDeviceSynchronize();
    return b;
}

void golb_destroy(golb_handle h)
{
    GolBatch *b = (GolBatch *)h;
    if (!b) return;
    cudaSetDevice(b->device);
    cudaFree(b->rows); cudaFree(b->frozen); cudaFree(b->frozen_val); cudaFree(b->obs);
    cudaFree(b->obs_u8); cudaFree(b->reward); cudaFree(b->pop); cudaFree(b->done);
    cudaFree(b->act_dev);
    if (b->h_act) cudaFreeHost(b->h_act);
    if (b->h_reward) cudaFreeHost(b->h_reward);
    if (b->ev) cudaEventDestroy(b->ev);
    free(b);
}

const char *golb_last_error(golb_handle h) { return h ? ((GolBatch *)h)->err : "null handle"; }

int golb_set_target(golb_handle h, float trg_pop, float range)
{
    GolBatch *b = (GolBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    // multi_env.py:146-160 set_params: plane = v / param_ranges[0]
    b->trg_pop = trg_pop;
    b->scalar_plane = trg_pop / range;
    b->max_loss = range;
    if (refresh_obs(b, 0)) return -1;
    CK(cudaDeviceSynchronize());       // the legacy stream is not ordered against a caller's non-blocking stream
    return 0;
}

int golb_reset(golb_handle h, uint32_t seed, float prob_life, void *stream)
{
    GolBatch *b = (GolBatch *)h;
    if (!b) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    int nrows = b->n_envs * b->width;
    uint32_t thresh = (uint32_t)((double)prob_life * 16777216.0);
    k_gol_populate<<<(nrows + 255) / 256, 256, 0, st>>>(b->rows, nrows, b->width, seed, thresh);
    CK(cudaGetLastError());
    if (b->classic) {
        CK(cudaMemsetAsync(b->frozen, 0, (size_t)nrows * 32, st));
        CK(cudaMemsetAsync(b->frozen_val, 0, (size_t)nrows * 32, st));
    }
    CK(cudaMemsetAsync(b->done, 0, b->n_envs, st));
    b->step_count = 0;
    return refresh_obs(b, st);
}

int golb_set_state(golb_handle h, const uint8_t *cells_host, void *stream)
{
    GolBatch *b = (GolBatch *)h;
    if (!b || !cells_host) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    size_t nrows = (size_t)b->n_envs * b->width, cells = nrows * b->width;
    uint8_t *tmp = NULL;
    CK(cudaMalloc(&tmp, cells));
    CK(cudaMemcpyAsync(tmp, cells_host, cells, cudaMemcpyHostToDevice, st));
    k_gol_pack<<<(int)((nrows + 255) / 256), 256, 0, st>>>(tmp, b->rows, (int)nrows, b->width);
    if (b->classic) {
        cudaMemsetAsync(b->frozen, 0, nrows * 32, st);
        cudaMemsetAsync(b->frozen_val, 0, nrows * 32, st);
    }
    int rc = refresh_obs(b, st);
    CK(cudaStreamSynchronize(st));
    cudaFree(tmp);
    b->step_count = 0;
    return rc;
}

int golb_get_state(golb_handle h, uint8_t *cells_host)
{
    GolBatch *b = (GolBatch *)h;
    if (!b || !cells_host) return -1;
    cudaSetDevice(b->device);
    size_t nrows = (size_t)b->n_envs * b->width;
    uint32_t *tmp = (uint32_t *)malloc(nrows * 4);
    CK(cudaDeviceSynchronize());
    cudaError_t e = cudaMemcpy(tmp, b->rows, nrows * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { free(tmp); return fail(b, "cudaMemcpy", e); }
    for (size_t i = 0; i < nrows; i++)
        for (int y = 0; y < b->width; y++) cells_host[i * b->width + y] = (tmp[i] >> y) & 1u;
    free(tmp);
    return 0;
}

static int gol_launch(GolBatch *b, const int64_t *actions_dev, int T, float *obs_seq, float *reward_seq, cudaStream_t st)
{
    const int w = b->width;
    float denom = b->max_loss * (float)b->max_step;       // torch: fp32 tensor * int, twice
    denom = denom * (float)b->n_envs;
    int threads = (256 / w) * w, epc = threads / w;
    int grid = (b->n_envs + epc - 1) / epc;
    if ((32 % w) == 0)
        k_gol_step<true><<<grid, threads, 0, st>>>(b->rows, actions_dev, b->obs, b->reward, b->pop, b->done, b->n_envs, w,
            b->trg_pop, b->max_loss, denom, b->step_count, b->max_step, T, obs_seq, reward_seq, b->scalar_plane);
    else
        k_gol_step<false><<<grid, threads, threads * 4, st>>>(b->rows, actions_dev, b->obs, b->reward, b->pop, b->done,
            b->n_envs, w, b->trg_pop, b->max_loss, denom, b->step_count, b->max_step, T, obs_seq, reward_seq, b->scalar_plane);
    CK(cudaGetLastError());
    b->step_count += T;
    return 0;
}

int golb_step(golb_handle h, const int64_t *actions_dev, void *stream)
{
    GolBatch *b = (GolBatch *)h;
    if (!b || !actions_dev) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int w = b->width;
    if (b->classic) {
        k_gol_classic_step<<<(b->n_envs + 63) / 64, 64, 0, st>>>(
            b->rows, b->frozen, b->frozen_val, actions_dev, b->obs_u8, b->reward, b->pop, b->done,
            b->n_envs, w, b->max_step, b->step_count);
        CK(cudaGetLastError());
        b->step_count++;
        return 0;
    }
    return gol_launch(b, actions_dev, 1, NULL, NULL, st);
}

int golb_step_n(golb_handle h, const int64_t *actions_dev, int T, float *obs_seq_dev, float *reward_seq_dev, void *stream)
{
    GolBatch *b = (GolBatch *)h;
    if (!b || !actions_dev || T < 1) return -1;
    if (b->classic) { snprintf(b->err, sizeof(b->err), "golb_step_n: multi-env handles only"); return -1; }
    cudaSetDevice(b->device);
    return gol_launch(b, actions_dev, T, obs_seq_dev, reward_seq_dev, (cudaStream_t)stream);
}

int golb_step_host(golb_handle h, const int64_t *actions_host, float *reward_host, void *stream)
{
    GolBatch *b = (GolBatch *)h;
    if (!b || !actions_host || !reward_host) return -1;
    cudaSetDevice(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(b->h_act, actions_host, (size_t)b->n_envs * 8);
    CK(cudaMemcpyAsync(b->act_dev, b->h_act, (size_t)b->n_envs * 8, cudaMemcpyHostToDevice, st));
    if (golb_step(h, b->act_dev, stream) != 0) return -1;
    CK(cudaMemcpyAsync(b->h_reward, b->reward, (size_t)b->n_envs * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(reward_host, b->h_reward, (size_t)b->n_envs * 4);
    return 0;
}

void *golb_ptr(golb_handle h, int which)
{
    GolBatch *b = (GolBatch *)h;
    if (!b) return NULL;
    switch (which) {
    case GOLB_ROWS: return b->rows;
    case GOLB_OBS: return b->classic ? (void *)b->obs_u8 : (void *)b->obs;
    case GOLB_REWARD: return b->reward;
    case GOLB_POP: return b->pop;
    case GOLB_DONE: return b->done;
    default: return NULL;
    }
}

int golb_step_count(golb_handle h) { return h ? ((GolBatch *)h)->step_count : -1; }

} // extern "C"
