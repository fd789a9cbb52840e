// Device-side audio-video multiblock mask sampling: the positions of every mask block of a
// batch are drawn ON THE GPU from a bit-exact replica of torch's CPU generator (MT19937, the reference's global RNG),
// the keep / drop index sets are built and compacted in the same launch, and no mask tensor ever crosses PCIe.
//
// Replaces the per-sample loop of the reference's collator (src/masks/avmultiblock3d.py:172-234: for every sample and
// every block `torch.randint` for (top, left, start) of the video block and (top, left) of the 4 x 6 audio block, boolean
// masks, `argwhere`, truncation to the batch minimum) for ALL mask generators of a step in one single-CTA launch.
// Draw order, the resample-on-empty-context rule and the ascending index order are the reference's, so with the same
// generator state the output is bit-identical to the host collator (tests/test_mask_collate_gpu.py); the generator state
// lives in device memory between steps and can be copied back into torch's host generator (same 624-word layout).
//
//   thread 0   : the sequential part -- MT19937 draws of one attempt (5 * npred values) into shared memory
//   all threads: keep flags of the 1568 video / 96 audio tokens, block-wide counts, exclusive scan, ordered compaction
//
// Work per step is tiny (B * generators attempts of a few microseconds on one SM, on a side stream); the point is the
// data path, not the arithmetic.
#include "common.cuh"

#define MC_THREADS 256
#define MC_MAX_BLOCKS 16

struct McGen {          // one mask generator of the step (block size already drawn by the host from the seeded per-step generator)
  int t, h, w;          // video block size in patches
  int npred;            // blocks per sample
  int max_ctx;          // max_context_duration (frames of the tubelet grid kept as context)
};

#define MC_MAX_GENS 8
struct McGens { McGen g[MC_MAX_GENS]; };   // kernel parameter (by value)

struct McState {        // torch's at::mt19937 data: state words, values left in the block, index of the next word
  uint32_t s[624];
  int left, next;
};

__device__ __forceinline__ void mc_next_state(McState* g) {
  uint32_t* st = g->s;
  const int N = 624, M = 397;
  auto tw = [](uint32_t u, uint32_t v) { return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u); };
  for (int j = 0; j < N - M; ++j) st[j] = st[j + M] ^ tw(st[j], st[j + 1]);
  for (int j = N - M; j < N - 1; ++j) st[j] = st[j + M - N] ^ tw(st[j], st[j + 1]);
  st[N - 1] = st[M - 1] ^ tw(st[N - 1], st[0]);
  g->left = 624;
  g->next = 0;
}
__device__ __forceinline__ uint32_t mc_rand32(McState* g) {
  if (--g->left == 0) mc_next_state(g);
  uint32_t y = g->s[g->next++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
// torch.randint(0, n, (1,)) on the CPU generator: random() % range for ranges below 2^32
__device__ __forceinline__ int mc_randint(McState* g, int n) { return (int)(mc_rand32(g) % (uint32_t)n); }

// Block-wide exclusive scan of one int per thread (MC_THREADS threads); returns the exclusive prefix, total in *total.
__device__ __forceinline__ int mc_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int base = 0;
  for (int i = 0; i < warp; ++i) base += warp_sums[i];
  int tot = 0;
  for (int i = 0; i < MC_THREADS / 32; ++i) tot += warp_sums[i];
  *total = tot;
  __syncthreads();
  return base + inc - v;
}

// Ordered compaction of `n` tokens: kept ids ascending into keep_out, dropped ids ascending into drop_out.
// flags[i] != 0 <=> token i is kept.  Returns the kept count.
__device__ __forceinline__ int mc_compact(const uint8_t* flags, int n, int64_t* keep_out, int64_t* drop_out, int* warp_sums) {
  const int per = (n + MC_THREADS - 1) / MC_THREADS;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += flags[i] ? 1 : 0;
  int total;
  int k = mc_scan(cnt, warp_sums, &total);
  int d = lo - k;
  for (int i = lo; i < hi; ++i) {
    if (flags[i]) keep_out[k++] = i;
    else drop_out[d++] = i;
  }
  return total;
}

__global__ void __launch_bounds__(MC_THREADS, 1)
mask_collate_kernel(McState* __restrict__ rng, const McGens gens, int n_gen, int B, int duration, int height, int width,
                    int a_height, int a_width, int a_bh, int a_bw, int64_t* __restrict__ enc_v, int64_t* __restrict__ pred_v,
                    int64_t* __restrict__ enc_a, int64_t* __restrict__ pred_a, int* __restrict__ counts, int* __restrict__ status) {
  __shared__ McState g;
  __shared__ int blocks[MC_MAX_BLOCKS][5];       // video (top, left, start), audio (top, left)
  __shared__ uint8_t keep_v[2048];
  __shared__ uint8_t keep_a[256];
  __shared__ int warp_sums[MC_THREADS / 32];
  __shared__ int n_keep_v_s;
  const int NV = duration * height * width, NA = a_height * a_width;
  for (int i = threadIdx.x; i < 624; i += MC_THREADS) g.s[i] = rng->s[i];
  if (threadIdx.x == 0) { g.left = rng->left; g.next = rng->next; }
  __syncthreads();
  int attempts = 0;
  for (int gi = 0; gi < n_gen; ++gi) {
    const McGen G = gens.g[gi];
    for (int s = 0; s < B;) {
      if (threadIdx.x == 0) {
        // reference draw order per block: video top, left, start, then audio top, left (avmultiblock3d.py:131-170)
        for (int k = 0; k < G.npred; ++k) {
          blocks[k][0] = mc_randint(&g, height - G.h + 1);
          blocks[k][1] = mc_randint(&g, width - G.w + 1);
          blocks[k][2] = mc_randint(&g, duration - G.t + 1);
          blocks[k][3] = mc_randint(&g, a_height - a_bh + 1);
          blocks[k][4] = mc_randint(&g, a_width - a_bw + 1);
        }
      }
      __syncthreads();
      int cnt = 0;
      for (int i = threadIdx.x; i < NV; i += MC_THREADS) {
        const int f = i / (height * width), r = (i / width) % height, c = i % width;
        bool keep = f < G.max_ctx || G.max_ctx >= duration;
        for (int k = 0; k < G.npred; ++k)
          keep = keep && !(f >= blocks[k][2] && f < blocks[k][2] + G.t && r >= blocks[k][0] && r < blocks[k][0] + G.h &&
                           c >= blocks[k][1] && c < blocks[k][1] + G.w);
        keep_v[i] = keep ? 1 : 0;
        cnt += keep ? 1 : 0;
      }
      for (int i = threadIdx.x; i < NA; i += MC_THREADS) {
        const int r = i / a_width, c = i % a_width;
        bool keep = true;
        for (int k = 0; k < G.npred; ++k)
          keep = keep && !(r >= blocks[k][3] && r < blocks[k][3] + a_bh && c >= blocks[k][4] && c < blocks[k][4] + a_bw);
        keep_a[i] = keep ? 1 : 0;
      }
      int total;
      mc_scan(cnt, warp_sums, &total);           // also a barrier: the flags are visible to everybody afterwards
      if (threadIdx.x == 0) n_keep_v_s = total;
      __syncthreads();
      ++attempts;
      if (n_keep_v_s == 0) {                       // empty video context: this sample is drawn again (:214-216)
        if (attempts > 64 * (B + 1) * n_gen) { if (threadIdx.x == 0) atomicOr(status, 2); break; }
        continue;
      }
      const int64_t row_v = ((int64_t)gi * B + s) * NV, row_a = ((int64_t)gi * B + s) * NA;
      const int kv = mc_compact(keep_v, NV, enc_v + row_v, pred_v + row_v, warp_sums);
      const int ka = mc_compact(keep_a, NA, enc_a + row_a, pred_a + row_a, warp_sums);
      if (threadIdx.x == 0) {
        int* c4 = counts + ((int64_t)gi * B + s) * 4;
        c4[0] = kv; c4[1] = ka; c4[2] = NV - kv; c4[3] = NA - ka;
        // the reference's len() of a squeezed one-element index set raises TypeError; the host collator reproduces that,
        // here it is reported through the status word (bit 0)
        if (kv == 1 || ka == 1 || NV - kv == 1 || NA - ka == 1) atomicOr(status, 1);
      }
      ++s;
      __syncthreads();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 624; i += MC_THREADS) rng->s[i] = g.s[i];
  if (threadIdx.x == 0) { rng->left = g.left; rng->next = g.next; }
}

extern "C" int avj_mask_collate(void* rng_state, const int32_t* gens, int n_gen, int B, int duration, int height, int width,
                                int a_height, int a_width, int a_block_h, int a_block_w, int64_t* enc_v, int64_t* pred_v,
                                int64_t* enc_a, int64_t* pred_a, int32_t* counts, int32_t* status, void* stream) {
  AVJ_CHECK(rng_state && gens && enc_v && pred_v && enc_a && pred_a && counts && status, "avj_mask_collate: NULL argument");
  AVJ_CHECK(n_gen >= 1 && n_gen <= MC_MAX_GENS && B >= 1, "avj_mask_collate: 1..%d generators and at least one sample", MC_MAX_GENS);
  McGens G;
  memset(&G, 0, sizeof(G));
  for (int i = 0; i < n_gen; ++i) {
    const int32_t* q = gens + 5 * i;                       // host array: t, h, w, npred, max_ctx per generator
    G.g[i].t = q[0]; G.g[i].h = q[1]; G.g[i].w = q[2]; G.g[i].npred = q[3]; G.g[i].max_ctx = q[4];
    AVJ_CHECK(q[3] >= 1 && q[3] <= MC_MAX_BLOCKS, "avj_mask_collate: 1..%d blocks per sample (got %d)", MC_MAX_BLOCKS, q[3]);
    AVJ_CHECK(q[0] >= 1 && q[0] <= duration && q[1] >= 1 && q[1] <= height && q[2] >= 1 && q[2] <= width && q[4] >= 1,
              "avj_mask_collate: block size %dx%dx%d does not fit the %dx%dx%d grid", q[0], q[1], q[2], duration, height, width);
  }
  AVJ_CHECK(a_block_h >= 1 && a_block_h <= a_height && a_block_w >= 1 && a_block_w <= a_width, "avj_mask_collate: audio block does not fit");
  AVJ_CHECK(duration * height * width <= 2048 && a_height * a_width <= 256, "avj_mask_collate: token grid too large (%d video, %d audio)",
            duration * height * width, a_height * a_width);
  mask_collate_kernel<<<1, MC_THREADS, 0, as_stream(stream)>>>(static_cast<McState*>(rng_state), G, n_gen, B,
                                                               duration, height, width, a_height, a_width, a_block_h, a_block_w, enc_v,
                                                               pred_v, enc_a, pred_a, counts, status);
  AVJ_LAUNCH_CHECK();
  return 0;
}
