// cc.cu — kernel 2: 3-D connected components on BIT ROWS (1 bit/voxel) with a lock-free
// union-find over x-runs, plus the mask morphology that follows it in meshify().
//
// What is reproduced (paths relative to /root/reference/):
//   src/bwlabel.c:478-543  bwlabel(): bubble fill = 6-connected labelling of the inverted mask, a
//                          background component is a bubble iff none of its voxels lies on one of
//                          the six volume faces (:498-529), skipped when <= 1 background component
//                          (:488-491); then the 18-connected foreground labelling
//   src/bwlabel.c:429-476  bwlabelCore(): keep the component with most voxels, ties to the lowest
//                          label = the component whose first voxel comes first in raster order
//   src/bwlabel.c:94-131, :154-245  the 18-neighbourhood: 6 faces + 12 edges (no corners)
//   src/meshify.c:218-264  dilate(): interior voxels only, 25 neighbours ((-1,-1,-1) never tested)
//   src/meshify.c:325-371  threshold, fill/largest application, edge darken, bright bounding box
//
// Design (nothing like the reference's raster scan + translation table): the mask is a bit row
// per (y,z); the unit of labelling is a "word run" = maximal run of set bits inside one 32-bit
// word.  Each word run owns the union-find slot  word*16 + (start_bit>>1)  (run starts inside a
// word are >= 2 bits apart, so slots are unique and ordered like the raster order of the runs'
// first voxels).  Unions are atomicMin-based (root = smallest slot of the component = the
// component's first voxel in raster order, which is the reference's tie-break for free).  Sizes
// and face-touch flags are accumulated per root; the selected components are written back as
// bit rows.  Labels per voxel are never materialised: CC traffic is ~0.4 B/voxel instead of the
// 16 B/voxel of a label volume.
#include "common.cuh"

struct cc_geom {
  int nx, ny, nz, w;
  long long nwords;
};

__device__ __forceinline__ uint32_t ld_parent(const uint2 *nodes, uint32_t a) {
  return __ldcg(reinterpret_cast<const unsigned int *>(&nodes[a].x));
}

__device__ __forceinline__ uint32_t uf_find(uint2 *nodes, uint32_t a) {
  uint32_t p = ld_parent(nodes, a);
  while (p != a) {
    uint32_t gp = ld_parent(nodes, p);
    if (gp != p) atomicMin(&nodes[a].x, gp);  // path splitting; monotone, so races are benign
    a = p;
    p = gp;
  }
  return a;
}

__device__ __forceinline__ void uf_union(uint2 *nodes, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf_find(nodes, a);
    b = uf_find(nodes, b);
    if (a == b) return;
    if (a < b) { uint32_t t = a; a = b; b = t; }
    uint32_t old = atomicMin(&nodes[a].x, b);  // a > b: hang a under b if a is still a root
    if (old == a) return;
    a = old;
  }
}

// start of the run of ones of `w` that contains bit b (bit b must be set)
__device__ __forceinline__ int run_start(uint32_t w, int b) {
  uint32_t below = (b == 31) ? 0xffffffffu : ((2u << b) - 1u);
  uint32_t z = ~w & below;
  return z ? 32 - __clz(z) : 0;
}
// last bit of the run of ones of `w` that starts at bit s
__device__ __forceinline__ int run_end(uint32_t w, int s) {
  uint32_t t = ~(w >> s);
  int len = t ? __ffs(t) - 1 : 32;
  return s + len - 1;
}
__device__ __forceinline__ uint32_t bits_range(int s, int e) {  // ones at s..e inclusive
  uint32_t hi = (e == 31) ? 0xffffffffu : ((2u << e) - 1u);
  return hi & ~((1u << s) - 1u);
}

__global__ void __launch_bounds__(256) k_cc_init(const uint32_t *__restrict__ bits, long long nwords,
                                                 uint2 *__restrict__ nodes) {
  long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (word >= nwords) return;
  uint32_t w = bits[word];
  uint32_t starts = w & ~(w << 1);
  while (starts) {
    int s = __ffs(starts) - 1;
    starts &= starts - 1;
    uint32_t slot = (uint32_t)word * 16u + (uint32_t)(s >> 1);
    nodes[slot] = make_uint2(slot, 0u);
  }
}

// union `me` with every run of the neighbour row word(s) that touches the run [s,e] of word xw.
// wide: also x-1 / x+1 (edge neighbours); else only the same x (face neighbour).
__device__ __forceinline__ void link_row(uint2 *nodes, uint32_t me, const uint32_t *__restrict__ nbrow,
                                         long long nbword0, int xw, int w, uint32_t runmask, int s, int e, bool wide) {
  uint32_t nw = __ldg(nbrow + xw);
  uint32_t m = runmask;
  if (wide) m |= (runmask << 1) | (runmask >> 1);
  uint32_t t = nw & m;
  while (t) {
    int b = __ffs(t) - 1;
    int st = run_start(nw, b);
    int en = run_end(nw, st);
    uf_union(nodes, me, (uint32_t)(nbword0 + xw) * 16u + (uint32_t)(st >> 1));
    t &= ~bits_range(st, en);
  }
  if (wide) {
    if (s == 0 && xw > 0) {
      uint32_t pw = __ldg(nbrow + xw - 1);
      if (pw >> 31) uf_union(nodes, me, (uint32_t)(nbword0 + xw - 1) * 16u + (uint32_t)(run_start(pw, 31) >> 1));
    }
    if (e == 31 && xw < w - 1) {
      uint32_t nx = __ldg(nbrow + xw + 1);
      if (nx & 1u) uf_union(nodes, me, (uint32_t)(nbword0 + xw + 1) * 16u);
    }
  }
}

// conn = 6 or 18.  One thread per bit word; backward neighbours only (each pair linked once).
__global__ void __launch_bounds__(256) k_cc_link(const uint32_t *__restrict__ bits, cc_geom g, uint2 *nodes, int conn) {
  long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (word >= g.nwords) return;
  uint32_t wv = __ldg(bits + word);
  if (!wv) return;
  long long row = word / g.w;
  int xw = (int)(word - row * g.w);
  int z = (int)(row / g.ny);
  int y = (int)(row - (long long)z * g.ny);
  const bool wide = conn >= 18;
  uint32_t rest = wv;
  while (rest) {
    int s = __ffs(rest) - 1;
    int e = run_end(wv, s);
    uint32_t rm = bits_range(s, e);
    rest &= ~rm;
    uint32_t me = (uint32_t)word * 16u + (uint32_t)(s >> 1);
    if (s == 0 && xw > 0) {  // the run continues from the previous word of this row
      uint32_t pw = __ldg(bits + word - 1);
      if (pw >> 31) uf_union(nodes, me, (uint32_t)(word - 1) * 16u + (uint32_t)(run_start(pw, 31) >> 1));
    }
    if (y > 0) link_row(nodes, me, bits + (row - 1) * g.w, (row - 1) * g.w, xw, g.w, rm, s, e, wide);
    if (z > 0) {
      long long r2 = row - g.ny;
      link_row(nodes, me, bits + r2 * g.w, r2 * g.w, xw, g.w, rm, s, e, wide);
      if (wide) {
        if (y > 0) link_row(nodes, me, bits + (r2 - 1) * g.w, (r2 - 1) * g.w, xw, g.w, rm, s, e, false);
        if (y < g.ny - 1) link_row(nodes, me, bits + (r2 + 1) * g.w, (r2 + 1) * g.w, xw, g.w, rm, s, e, false);
      }
    }
  }
}

__device__ __forceinline__ void flush_stats(uint2 *nodes, uint32_t root, uint32_t cnt, uint32_t flag) {
  // warp-aggregated: lanes holding the same root combine before touching memory
  unsigned peers = __match_any_sync(__activemask(), root);
  uint32_t tot = __reduce_add_sync(peers, cnt);
  uint32_t fl = __reduce_or_sync(peers, flag);
  if ((unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u) && root != 0xffffffffu) {
    if (tot) atomicAdd(&nodes[root].y, tot);
    if (fl) atomicOr(&nodes[root].y, 0x80000000u);
  }
}

// flatten every word run to its root; accumulate voxel counts (low 31 bits of node.y) and the
// "touches a volume face" flag (bit 31) per root.
__global__ void __launch_bounds__(256) k_cc_flatten(const uint32_t *__restrict__ bits, cc_geom g, uint2 *nodes) {
  long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t wv = word < g.nwords ? __ldg(bits + word) : 0u;
  long long row = word < g.nwords ? word / g.w : 0;
  int xw = (int)(word - row * g.w);
  int z = (int)(row / g.ny);
  int y = (int)(row - (long long)z * g.ny);
  const bool rowface = (y == 0) || (y == g.ny - 1) || (z == 0) || (z == g.nz - 1);
  uint32_t curRoot = 0xffffffffu, curCnt = 0, curFlag = 0;
  uint32_t rest = wv;
  // all lanes iterate together until every lane is out of runs (keeps the match/reduce converged)
  while (__any_sync(0xffffffffu, rest != 0)) {
    uint32_t root = 0xffffffffu, cnt = 0, flag = 0;
    if (rest) {
      int s = __ffs(rest) - 1;
      int e = run_end(wv, s);
      uint32_t rm = bits_range(s, e);
      rest &= ~rm;
      uint32_t slot = (uint32_t)word * 16u + (uint32_t)(s >> 1);
      root = uf_find(nodes, slot);
      if (root != slot) atomicMin(&nodes[slot].x, root);
      cnt = (uint32_t)(e - s + 1);
      int x0 = xw * 32 + s, x1 = xw * 32 + e;
      flag = (rowface || x0 == 0 || x1 == g.nx - 1) ? 1u : 0u;
    }
    if (root == curRoot) {
      curCnt += cnt;
      curFlag |= flag;
      root = 0xffffffffu; cnt = 0; flag = 0;
    }
    // flush the previous accumulator of lanes whose root changed
    bool changed = root != 0xffffffffu;
    uint32_t fr = changed ? curRoot : 0xffffffffu, fc = changed ? curCnt : 0, ff = changed ? curFlag : 0;
    if (__any_sync(0xffffffffu, fr != 0xffffffffu)) flush_stats(nodes, fr, fc, ff);
    if (changed) { curRoot = root; curCnt = cnt; curFlag = flag; }
  }
  flush_stats(nodes, curRoot, curCnt, curFlag);
}

// number of components and the largest one: key = (size << 32) | ~rootslot, so that among equal
// sizes the smallest slot (earliest first voxel in raster order) wins, as src/bwlabel.c:462-466.
__global__ void __launch_bounds__(256) k_cc_best(const uint32_t *__restrict__ bits, long long nwords,
                                                 const uint2 *__restrict__ nodes, unsigned long long *best,
                                                 unsigned int *nroots) {
  long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t w = word < nwords ? __ldg(bits + word) : 0u;
  uint32_t starts = w & ~(w << 1);
  unsigned long long key = 0;
  unsigned int cnt = 0;
  while (starts) {
    int s = __ffs(starts) - 1;
    starts &= starts - 1;
    uint32_t slot = (uint32_t)word * 16u + (uint32_t)(s >> 1);
    uint2 nd = nodes[slot];
    if (nd.x == slot) {
      cnt++;
      unsigned long long k = ((unsigned long long)(nd.y & 0x7fffffffu) << 32) | (unsigned long long)(0xffffffffu - slot);
      key = k > key ? k : key;
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, key, d);
    key = o > key ? o : key;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (key && best) atomicMax(best, key);
    if (cnt) atomicAdd(nroots, cnt);
  }
}

// mode 0: out = runs of `bits` whose root is the largest component
// mode 1: out = other | runs of `bits` (the background) whose root does not touch a face, only if
//         there are >= 2 background components (src/bwlabel.c:488-491); else out = other
__global__ void __launch_bounds__(256) k_cc_select(const uint32_t *__restrict__ bits, long long nwords,
                                                   const uint2 *__restrict__ nodes, int mode,
                                                   const unsigned long long *__restrict__ best,
                                                   const unsigned int *__restrict__ nroots,
                                                   const uint32_t *__restrict__ other, uint32_t *__restrict__ out) {
  long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (word >= nwords) return;
  uint32_t wv = __ldg(bits + word);
  uint32_t res = 0;
  if (mode == 0) {
    uint32_t bestslot = 0xffffffffu - (uint32_t)(*best & 0xffffffffull);
    bool have = *best != 0ull;
    uint32_t rest = wv;
    while (rest && have) {
      int s = __ffs(rest) - 1;
      int e = run_end(wv, s);
      uint32_t rm = bits_range(s, e);
      rest &= ~rm;
      if (nodes[(uint32_t)word * 16u + (uint32_t)(s >> 1)].x == bestslot) res |= rm;
    }
  } else {
    res = __ldg(other + word);
    if (*nroots > 1u) {
      uint32_t rest = wv;
      while (rest) {
        int s = __ffs(rest) - 1;
        int e = run_end(wv, s);
        uint32_t rm = bits_range(s, e);
        rest &= ~rm;
        uint32_t root = nodes[(uint32_t)word * 16u + (uint32_t)(s >> 1)].x;
        if (!(nodes[root].y >> 31)) res |= rm;
      }
    }
  }
  out[word] = res;
}

// keep = largest | dilate25(largest) (interior voxels only), and the bounding box of the bright
// voxels  bright = fillOrFg & keep  (keep == all ones when largest == nullptr).
__global__ void __launch_bounds__(256) k_dilate_bbox(const uint32_t *__restrict__ largest,
                                                     const uint32_t *__restrict__ bright_src, cc_geom g,
                                                     uint32_t *__restrict__ keep, int *__restrict__ lohi) {
  long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int lo0 = INT_MAX, lo1 = INT_MAX, lo2 = INT_MAX, hi0 = -1, hi1 = -1, hi2 = -1;
  if (word < g.nwords) {
    long long row = word / g.w;
    int xw = (int)(word - row * g.w);
    int z = (int)(row / g.ny);
    int y = (int)(row - (long long)z * g.ny);
    uint32_t k = 0xffffffffu;
    if (largest) {
      k = __ldg(largest + word);
      if (y >= 1 && y <= g.ny - 2 && z >= 1 && z <= g.nz - 2) {
        uint32_t acc = 0;
#pragma unroll
        for (int dz = -1; dz <= 1; dz++)
#pragma unroll
          for (int dy = -1; dy <= 1; dy++) {
            const uint32_t *r = largest + (row + dy + (long long)dz * g.ny) * g.w;
            uint32_t c = __ldg(r + xw);
            uint32_t p = xw > 0 ? __ldg(r + xw - 1) : 0u;
            uint32_t n = xw < g.w - 1 ? __ldg(r + xw + 1) : 0u;
            uint32_t left = (c << 1) | (p >> 31);   // bit x = voxel x-1  (neighbour dx = -1)
            uint32_t right = (c >> 1) | (n << 31);  // bit x = voxel x+1  (neighbour dx = +1)
            acc |= c | right;
            if (!(dz == -1 && dy == -1)) acc |= left;  // (-1,-1,-1) is never tested (meshify.c:252)
          }
        // interior x only: 1 .. nx-2
        int xb = xw * 32;
        uint32_t im = 0xffffffffu;
        if (xb == 0) im &= ~1u;
        int last = g.nx - 2 - xb;  // highest interior bit in this word
        if (last < 0) im = 0;
        else if (last < 31) im &= (2u << last) - 1u;
        k |= acc & im;
      }
      keep[word] = k;
    }
    uint32_t b = __ldg(bright_src + word) & k;
    if (b) {
      lo0 = xw * 32 + __ffs(b) - 1;
      hi0 = xw * 32 + 31 - __clz(b);
      lo1 = hi1 = y;
      lo2 = hi2 = z;
    }
  }
  lo0 = __reduce_min_sync(0xffffffffu, lo0); lo1 = __reduce_min_sync(0xffffffffu, lo1); lo2 = __reduce_min_sync(0xffffffffu, lo2);
  hi0 = __reduce_max_sync(0xffffffffu, hi0); hi1 = __reduce_max_sync(0xffffffffu, hi1); hi2 = __reduce_max_sync(0xffffffffu, hi2);
  if ((threadIdx.x & 31) == 0 && hi0 >= 0) {
    atomicMin(&lohi[0], lo0); atomicMin(&lohi[1], lo1); atomicMin(&lohi[2], lo2);
    atomicMax(&lohi[3], hi0); atomicMax(&lohi[4], hi1); atomicMax(&lohi[5], hi2);
  }
}

// ------------------------------------------------------------------------------------------------
// materialise the composed volume / mask (parity hooks) and, lazily, its minimum
__global__ void __launch_bounds__(256) k_compose(compose_params c, float *__restrict__ out, uint8_t *__restrict__ mask,
                                                 unsigned int *__restrict__ cmin_enc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)c.nx * c.ny * c.nz;
  float v = INFINITY;
  if (i < n) {
    int x = (int)(i % c.nx);
    size_t r = i / c.nx;
    int y = (int)(r % c.ny), z = (int)(r / c.ny);
    v = composed_value(c, x, y, z);
    if (out) out[i] = v;
    if (mask) {
      const uint32_t *m = c.keep ? c.keep : c.fill;
      mask[i] = m ? (uint8_t)((m[(r * c.w) + (x >> 5)] >> (x & 31)) & 1u) : 0;
    }
  }
  if (cmin_enc) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0 && v != INFINITY) atomicMin(cmin_enc, f32_enc(v));
  }
}

int b2m_compose_materialize(b2m_ctx *ctx, const b2m_geom &g, const b2m_front_out *fo, float *d_composed,
                            uint8_t *d_mask, b2m_scalars *d_sc, int want_min) {
  compose_params c;
  c.S = fo->S; c.fill = fo->fill; c.keep = fo->keep;
  c.nx = g.nx; c.ny = g.ny; c.nz = g.nz; c.w = g.w;
  c.iso = fo->iso; c.mn = fo->vmin; c.edge_max = fo->edge_max;
  KT_LAUNCH(ctx, "compose", k_compose<<<b2m_cdiv(g.n, 256), 256, 0, ctx->stream>>>(c, d_composed, d_mask, want_min ? &d_sc->cmin_enc : nullptr));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

static int cc_label(b2m_ctx *ctx, const uint32_t *bits, const cc_geom &cg, uint2 *nodes, int conn) {
  unsigned blocks = b2m_cdiv(cg.nwords, 256);
  KT_LAUNCH(ctx, "cc_init", k_cc_init<<<blocks, 256, 0, ctx->stream>>>(bits, cg.nwords, nodes));
  KT_LAUNCH(ctx, "cc_link", k_cc_link<<<blocks, 256, 0, ctx->stream>>>(bits, cg, nodes, conn));
  KT_LAUNCH(ctx, "cc_flatten", k_cc_flatten<<<blocks, 256, 0, ctx->stream>>>(bits, cg, nodes));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

// CC part of the front: fills fo->fill / fo->keep and the raw bright bbox in d_sc->lo/hi.
int b2m_cc_run(b2m_ctx *ctx, const b2m_geom &g, const b2m_opts *o, b2m_scalars *d_sc, b2m_front_out *fo) {
  cc_geom cg = {g.nx, g.ny, g.nz, g.w, g.nwords};
  size_t wbytes = (size_t)g.nwords * 4;
  unsigned blocks = b2m_cdiv(g.nwords, 256);
  const bool cc = o->only_largest || o->fill_bubbles;
  B2M_TRY(b2m_reserve(ctx, BUF_FG, wbytes));
  uint32_t *fg = b2m_ptr<uint32_t>(ctx, BUF_FG);
  uint32_t *bg = nullptr;
  if (o->fill_bubbles) {
    B2M_TRY(b2m_reserve(ctx, BUF_BG, wbytes));
    bg = b2m_ptr<uint32_t>(ctx, BUF_BG);
  }
  B2M_TRY(b2m_threshold_run(ctx, fo->S, g, fo->iso, fg, bg));
  fo->fill = nullptr;
  fo->keep = nullptr;
  const uint32_t *bright = fg;
  uint2 *nodes = nullptr;
  if (cc) {
    if ((unsigned long long)g.nwords * 16ull > 0xffffffffull) {
      b2m_set_error("volume too large for 32-bit run slots (%lld words)", g.nwords);
      return B2M_EARG;
    }
    B2M_TRY(b2m_reserve(ctx, BUF_NODES, (size_t)g.nwords * 16 * sizeof(uint2)));
    nodes = b2m_ptr<uint2>(ctx, BUF_NODES);
  }
  if (o->fill_bubbles) {
    B2M_TRY(b2m_reserve(ctx, BUF_FILL, wbytes));
    uint32_t *fill = b2m_ptr<uint32_t>(ctx, BUF_FILL);
    B2M_TRY(cc_label(ctx, bg, cg, nodes, 6));
    KT_LAUNCH(ctx, "cc_best", k_cc_best<<<blocks, 256, 0, ctx->stream>>>(bg, g.nwords, nodes, nullptr, &d_sc->nroots_bg));
    KT_LAUNCH(ctx, "cc_select", k_cc_select<<<blocks, 256, 0, ctx->stream>>>(bg, g.nwords, nodes, 1, nullptr, &d_sc->nroots_bg, fg, fill));
    fo->fill = fill;
    bright = fill;
  }
  uint32_t *largest = nullptr, *keep = nullptr;
  if (o->only_largest) {
    B2M_TRY(b2m_reserve(ctx, BUF_LARGEST, wbytes));
    B2M_TRY(b2m_reserve(ctx, BUF_KEEP, wbytes));
    largest = b2m_ptr<uint32_t>(ctx, BUF_LARGEST);
    keep = b2m_ptr<uint32_t>(ctx, BUF_KEEP);
    B2M_TRY(cc_label(ctx, bright, cg, nodes, 18));
    KT_LAUNCH(ctx, "cc_best", k_cc_best<<<blocks, 256, 0, ctx->stream>>>(bright, g.nwords, nodes, &d_sc->best_fg, &d_sc->nroots_fg));
    KT_LAUNCH(ctx, "cc_select", k_cc_select<<<blocks, 256, 0, ctx->stream>>>(bright, g.nwords, nodes, 0, &d_sc->best_fg, nullptr, nullptr, largest));
    fo->keep = keep;
  }
  KT_LAUNCH(ctx, "dilate_bbox", k_dilate_bbox<<<blocks, 256, 0, ctx->stream>>>(largest, bright, cg, keep, d_sc->lo));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}
