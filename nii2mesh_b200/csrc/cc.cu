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

// ---- tile-local labelling ----------------------------------------------------------------------
// A CTA labels a tile of CT_W x CT_Y x CT_Z bit words (256 x 8 x 8 voxels) entirely in shared memory
// (same run slots, same union-by-minimum rule, ~30-cycle shared-memory hops instead of L2 round
// trips), then publishes one global node per run: parent = the tile-local root's GLOBAL slot, and on
// the local roots the voxel count and face flag of the whole local component.  Only neighbour pairs
// that straddle two tiles are left for the global union-find (k_cc_border), which therefore works
// on a forest of a few tile-roots per tile instead of one node per run.
#define CT_W 8
#define CT_Y 8
#define CT_Z 8
#define CT_WORDS (CT_W * CT_Y * CT_Z)

// shared-memory entry of a run slot: parent local slot << 16 | face flag << 15 | voxel count (<= 16384)
__device__ __forceinline__ uint32_t lfind(volatile uint32_t *par, uint32_t a) {
  uint32_t p = par[a] >> 16;
  while (p != a) {
    a = p;
    p = par[a] >> 16;
  }
  return a;
}
// find with path halving; only valid while the count bits are still zero (phase A)
__device__ __forceinline__ uint32_t lfind_compress(uint32_t *par, uint32_t a) {
  uint32_t p = ((volatile uint32_t *)par)[a] >> 16;
  while (p != a) {
    const uint32_t gp = ((volatile uint32_t *)par)[p] >> 16;
    if (gp != p) atomicMin(&par[a], gp << 16);
    a = p;
    p = gp;
  }
  return a;
}
__device__ __forceinline__ void lunion(uint32_t *par, uint32_t a, uint32_t b) {
  for (;;) {
    a = lfind_compress(par, a);
    b = lfind_compress(par, b);
    if (a == b) return;
    if (a < b) { uint32_t t = a; a = b; b = t; }
    uint32_t old = atomicMin(&par[a], b << 16) >> 16;  // counts are still zero in this phase
    if (old == a) return;
    a = old;
  }
}

// Enumerates the backward neighbour runs of run [s,e] (mask rm) of the word at tile-local (lx,ly,lz)
// and calls link(neighbour word delta (dx,dy,dz), neighbour word bits, start bit of the neighbour run).
// fetch(dx,dy,dz) returns the neighbour word's bits or 0 when that pair is not this phase's business.
template <int CONN, class Fetch, class Link>
__device__ __forceinline__ void cc_visit_neighbours(uint32_t rm, int s, int e, Fetch fetch, Link link) {
  auto row = [&](int dy, int dz, bool wide) {
    const uint32_t nw = fetch(0, dy, dz);
    uint32_t m = rm;
    if (wide) m |= (rm << 1) | (rm >> 1);
    uint32_t t = nw & m;
    while (t) {
      const int b = __ffs(t) - 1;
      const int st = run_start(nw, b);
      const int en = run_end(nw, st);
      link(0, dy, dz, st);
      t &= ~bits_range(st, en);
    }
    if (wide) {
      if (s == 0) {
        const uint32_t pw = fetch(-1, dy, dz);
        if (pw >> 31) link(-1, dy, dz, run_start(pw, 31));
      }
      if (e == 31) {
        const uint32_t nx = fetch(1, dy, dz);
        if (nx & 1u) link(1, dy, dz, 0);
      }
    }
  };
  if (s == 0) {  // the run continues from the previous word of this row
    const uint32_t pw = fetch(-1, 0, 0);
    if (pw >> 31) link(-1, 0, 0, run_start(pw, 31));
  }
  const bool wide = CONN >= 18;
  row(-1, 0, wide);
  row(0, -1, wide);
  if (wide) {
    row(-1, -1, false);
    row(1, -1, false);
  }
}

template <int CONN>
__global__ void __launch_bounds__(CT_WORDS) k_cc_local(const uint32_t *__restrict__ bits, cc_geom g, uint2 *__restrict__ nodes) {
  __shared__ uint32_t sb[CT_WORDS];
  __shared__ uint32_t par[CT_WORDS * 16];
  const int t = threadIdx.x;
  const int lx = t % CT_W, ly = (t / CT_W) % CT_Y, lz = t / (CT_W * CT_Y);
  const int xw = blockIdx.x * CT_W + lx, y = blockIdx.y * CT_Y + ly, z = blockIdx.z * CT_Z + lz;
  const bool valid = xw < g.w && y < g.ny && z < g.nz;
  const long long word = valid ? ((long long)z * g.ny + y) * g.w + xw : 0;
  const uint32_t wv = valid ? __ldg(bits + word) : 0u;
  sb[t] = wv;
  {
    uint32_t starts = wv & ~(wv << 1);
    while (starts) {
      const int s = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint32_t slot = (uint32_t)t * 16u + (uint32_t)(s >> 1);
      par[slot] = slot << 16;
    }
  }
  __syncthreads();
  auto fetch = [&](int dx, int dy, int dz) -> uint32_t {
    const int ax = lx + dx, ay = ly + dy, az = lz + dz;
    if ((unsigned)ax >= CT_W || (unsigned)ay >= CT_Y || (unsigned)az >= CT_Z) return 0u;  // other tile: k_cc_border
    return sb[(az * CT_Y + ay) * CT_W + ax];
  };
  // phase A: unions inside the tile
  for (uint32_t rest = wv; rest;) {
    const int s = __ffs(rest) - 1;
    const int e = run_end(wv, s);
    const uint32_t rm = bits_range(s, e);
    rest &= ~rm;
    const uint32_t me = (uint32_t)t * 16u + (uint32_t)(s >> 1);
    cc_visit_neighbours<CONN>(rm, s, e, fetch, [&](int dx, int dy, int dz, int st) {
      const int nt = ((lz + dz) * CT_Y + (ly + dy)) * CT_W + (lx + dx);
      lunion(par, me, (uint32_t)nt * 16u + (uint32_t)(st >> 1));
    });
  }
  __syncthreads();
  // phase B: local roots collect the voxel count and the face flag of their local component.
  // Lanes that reach the same root combine first (one shared-memory atomic per warp and root: 512
  // same-address atomics per tile were the bottleneck of the first version).
  const bool rowface = (y == 0) || (y == g.ny - 1) || (z == 0) || (z == g.nz - 1);
  uint32_t myroot[16];  // local root per run of this word, in run order (<= 16 runs)
  int nrun = 0;
  {
    uint32_t rest = wv;
    while (__any_sync(0xffffffffu, rest != 0)) {
      uint32_t r = 0xffffffffu, cnt = 0, flag = 0;
      if (rest) {
        const int s = __ffs(rest) - 1;
        const int e = run_end(wv, s);
        rest &= ~bits_range(s, e);
        r = lfind(par, (uint32_t)t * 16u + (uint32_t)(s >> 1));
        myroot[nrun++] = r;
        const int x0 = xw * 32 + s, x1 = xw * 32 + e;
        cnt = (uint32_t)(e - s + 1);
        flag = (rowface || x0 == 0 || x1 == g.nx - 1) ? 0x8000u : 0u;
      }
      const unsigned peers = __match_any_sync(0xffffffffu, r);
      const uint32_t tot = __reduce_add_sync(peers, cnt);
      const uint32_t fl = __reduce_or_sync(peers, flag);
      if (r != 0xffffffffu && (unsigned)(__ffs(peers) - 1) == (unsigned)(t & 31)) {
        atomicAdd(&par[r], tot);
        if (fl) atomicOr(&par[r], fl);
      }
    }
  }
  __syncthreads();
  // phase C: publish the global nodes
  const int tx0 = blockIdx.x * CT_W, ty0 = blockIdx.y * CT_Y, tz0 = blockIdx.z * CT_Z;
  nrun = 0;
  for (uint32_t rest = wv; rest; nrun++) {
    const int s = __ffs(rest) - 1;
    const int e = run_end(wv, s);
    rest &= ~bits_range(s, e);
    const uint32_t me = (uint32_t)t * 16u + (uint32_t)(s >> 1);
    const uint32_t r = myroot[nrun];
    const int rt = (int)(r >> 4);
    const int rx = rt % CT_W, ry = (rt / CT_W) % CT_Y, rz = rt / (CT_W * CT_Y);
    const long long rword = ((long long)(tz0 + rz) * g.ny + (ty0 + ry)) * g.w + (tx0 + rx);
    const uint32_t gparent = (uint32_t)rword * 16u + (r & 15u);
    uint32_t stat = 0;
    if (r == me) {
      const uint32_t pe = par[me];
      stat = (pe & 0x7fffu) | ((pe & 0x8000u) << 16);
    }
    nodes[(uint32_t)word * 16u + (uint32_t)(s >> 1)] = make_uint2(gparent, stat);
  }
}

// neighbour pairs that straddle two tiles: global lock-free unions between (mostly) tile roots
template <int CONN>
__global__ void __launch_bounds__(256) k_cc_border(const uint32_t *__restrict__ bits, cc_geom g, uint2 *nodes) {
  const long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (word >= g.nwords) return;
  const long long rowi = word / g.w;
  const int xw = (int)(word - rowi * g.w);
  const int z = (int)(rowi / g.ny);
  const int y = (int)(rowi - (long long)z * g.ny);
  const int lx = xw % CT_W, ly = y % CT_Y, lz = z % CT_Z;
  // only words on a tile face can have a backward neighbour in another tile
  if (!(lx == 0 || lx == CT_W - 1 || ly == 0 || ly == CT_Y - 1 || lz == 0)) return;
  const uint32_t wv = __ldg(bits + word);
  if (!wv) return;
  auto fetch = [&](int dx, int dy, int dz) -> uint32_t {
    const int ax = lx + dx, ay = ly + dy, az = lz + dz;
    if ((unsigned)ax < CT_W && (unsigned)ay < CT_Y && (unsigned)az < CT_Z) return 0u;  // same tile: done locally
    const int gx = xw + dx, gy = y + dy, gz = z + dz;
    if (gx < 0 || gx >= g.w || gy < 0 || gy >= g.ny || gz < 0) return 0u;
    return __ldg(bits + ((long long)gz * g.ny + gy) * g.w + gx);
  };
  for (uint32_t rest = wv; rest;) {
    const int s = __ffs(rest) - 1;
    const int e = run_end(wv, s);
    const uint32_t rm = bits_range(s, e);
    rest &= ~rm;
    const uint32_t me = (uint32_t)word * 16u + (uint32_t)(s >> 1);
    cc_visit_neighbours<CONN>(rm, s, e, fetch, [&](int dx, int dy, int dz, int st) {
      const long long nword = ((long long)(z + dz) * g.ny + (y + dy)) * g.w + (xw + dx);
      uf_union(nodes, me, (uint32_t)nword * 16u + (uint32_t)(st >> 1));
    });
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

// every tile root that lost its root status in k_cc_border hands the count / face flag of its local
// component to its final root and is pointed straight at it (runs then reach the final root in two hops).
__global__ void __launch_bounds__(256) k_cc_flatten(const uint32_t *__restrict__ bits, cc_geom g, uint2 *nodes) {
  long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t wv = word < g.nwords ? __ldg(bits + word) : 0u;
  uint32_t starts = wv & ~(wv << 1);
  // all lanes iterate together until every lane is out of runs (keeps the match/reduce converged)
  while (__any_sync(0xffffffffu, starts != 0)) {
    uint32_t root = 0xffffffffu, cnt = 0, flag = 0;
    if (starts) {
      const int s = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint32_t slot = (uint32_t)word * 16u + (uint32_t)(s >> 1);
      const uint2 nd = nodes[slot];
      if ((nd.y & 0x7fffffffu) && nd.x != slot) {  // a tile root (it owns a count) that is no longer a root
        root = uf_find(nodes, slot);
        atomicMin(&nodes[slot].x, root);
        cnt = nd.y & 0x7fffffffu;
        flag = nd.y >> 31;
      }
    }
    if (__any_sync(0xffffffffu, root != 0xffffffffu)) flush_stats(nodes, root, cnt, flag);
  }
}

// final root of a run after k_cc_flatten: run -> tile root -> final root
__device__ __forceinline__ uint32_t cc_final_root(const uint2 *__restrict__ nodes, uint32_t slot) {
  uint32_t p = nodes[slot].x;
  while (p != slot) {
    slot = p;
    p = nodes[slot].x;
  }
  return slot;
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
      if (cc_final_root(nodes, (uint32_t)word * 16u + (uint32_t)(s >> 1)) == bestslot) res |= rm;
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
        uint32_t root = cc_final_root(nodes, (uint32_t)word * 16u + (uint32_t)(s >> 1));
        if (!(nodes[root].y >> 31)) res |= rm;
      }
    }
  }
  out[word] = res;
}

// keep = largest | dilate25(largest) (interior voxels only), and the bounding box of the bright
// voxels  bright = fillOrFg & keep  (keep == all ones when largest == nullptr).
// A thread owns one bit-word column (xw, y) and marches along z: per new plane it loads the three
// rows y-1, y, y+1 once (x neighbours by funnel shifts with the adjacent words) and keeps
//   PF(z) = OR over dy of full(y+dy, z),  full = c | c<<1 | c>>1 (with carries)    [all 9 in-plane taps]
//   Q(z)  = full(y,z) | full(y+1,z) | nol(y-1,z),  nol = c | c>>1                  [plane z-1 taps: the
//           reference's dilate() never tests offset (-1,-1,-1), src/meshify.c:252]
// so that out(z) = Q(z-1) | PF(z) | PF(z+1): 3 word loads per output word instead of 27.
#define DIL_ZC 16
__global__ void __launch_bounds__(256) k_dilate_bbox(const uint32_t *__restrict__ largest,
                                                     const uint32_t *__restrict__ bright_src, cc_geom g,
                                                     uint32_t *__restrict__ keep, int *__restrict__ lohi) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (y, xw) flattened
  const long long ncol = (long long)g.ny * g.w;
  int lo0 = INT_MAX, lo1 = INT_MAX, lo2 = INT_MAX, hi0 = -1, hi1 = -1, hi2 = -1;
  if (col < ncol) {
    const int y = (int)(col / g.w), xw = (int)(col - (long long)y * g.w);
    const int z0 = blockIdx.y * DIL_ZC, z1 = min(z0 + DIL_ZC, g.nz);
    const long long plane = ncol;
    // interior x mask of this word: voxels 1 .. nx-2
    uint32_t im = 0xffffffffu;
    {
      const int xb = xw * 32;
      if (xb == 0) im &= ~1u;
      const int last = g.nx - 2 - xb;  // highest interior bit in this word
      if (last < 0) im = 0;
      else if (last < 31) im &= (2u << last) - 1u;
    }
    const bool yin = y >= 1 && y <= g.ny - 2;
    uint32_t Qm1 = 0, PF0 = 0, Q0 = 0, c0 = 0;  // Q(z-1), PF(z), Q(z), largest word at z
    auto load_plane = [&](int z, uint32_t &PF, uint32_t &Q, uint32_t &cc) {
      PF = Q = cc = 0;
      if (z < 0 || z >= g.nz) return;
      const uint32_t *base = largest + (long long)z * plane + col;
#pragma unroll
      for (int dy = -1; dy <= 1; dy++) {
        if (y + dy < 0 || y + dy >= g.ny) continue;
        const uint32_t *r = base + (long long)dy * g.w;
        const uint32_t c = __ldg(r);
        const uint32_t pw = xw > 0 ? __ldg(r - 1) : 0u;
        const uint32_t nw = xw < g.w - 1 ? __ldg(r + 1) : 0u;
        const uint32_t left = (c << 1) | (pw >> 31);   // bit x = voxel x-1
        const uint32_t right = (c >> 1) | (nw << 31);  // bit x = voxel x+1
        PF |= c | left | right;
        Q |= (dy == -1) ? (c | right) : (c | left | right);
        if (dy == 0) cc = c;
      }
    };
    if (largest) {
      uint32_t t0, t1;
      load_plane(z0 - 1, t0, Qm1, t1);
      load_plane(z0, PF0, Q0, c0);
    }
    for (int z = z0; z < z1; z++) {
      const long long word = (long long)z * plane + col;
      uint32_t k = 0xffffffffu;
      if (largest) {
        uint32_t PF1, Q1, c1;
        load_plane(z + 1, PF1, Q1, c1);
        k = c0;
        if (yin && z >= 1 && z <= g.nz - 2) k |= (Qm1 | PF0 | PF1) & im;
        keep[word] = k;
        Qm1 = Q0; PF0 = PF1; Q0 = Q1; c0 = c1;
      }
      const uint32_t bb = __ldg(bright_src + word) & k;
      if (bb) {
        lo0 = min(lo0, xw * 32 + __ffs(bb) - 1);
        hi0 = max(hi0, xw * 32 + 31 - __clz(bb));
        lo1 = min(lo1, y); hi1 = max(hi1, y);
        lo2 = min(lo2, z); hi2 = max(hi2, z);
      }
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
  dim3 tiles(b2m_cdiv(cg.w, CT_W), b2m_cdiv(cg.ny, CT_Y), b2m_cdiv(cg.nz, CT_Z));
  if (conn >= 18) {
    KT_LAUNCH(ctx, "cc_local", k_cc_local<18><<<tiles, CT_WORDS, 0, ctx->stream>>>(bits, cg, nodes));
    KT_LAUNCH(ctx, "cc_border", k_cc_border<18><<<blocks, 256, 0, ctx->stream>>>(bits, cg, nodes));
  } else {
    KT_LAUNCH(ctx, "cc_local", k_cc_local<6><<<tiles, CT_WORDS, 0, ctx->stream>>>(bits, cg, nodes));
    KT_LAUNCH(ctx, "cc_border", k_cc_border<6><<<blocks, 256, 0, ctx->stream>>>(bits, cg, nodes));
  }
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
  {
    dim3 dgrid(b2m_cdiv((size_t)g.ny * g.w, 256), b2m_cdiv(g.nz, DIL_ZC));
    KT_LAUNCH(ctx, "dilate_bbox", k_dilate_bbox<<<dgrid, 256, 0, ctx->stream>>>(largest, bright, cg, keep, d_sc->lo));
  }
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}
