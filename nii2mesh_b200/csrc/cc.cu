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
  int zface_lo, zface_hi;  // plane 0 / plane nz-1 of the labelled region is a face of the whole volume (slabs)
};
// Run slots stay below 2^31 (a slab holds at most 2^27 bit words = 2^32 voxels), so a parent field with bit 31 set is
// never a slot: it marks a seam root (slabs) - transiently CC_TAG | ticket while the seams are being resolved, and
// CC_SENT once the root is known to belong to the selected (largest) component.
#define CC_TAG 0x80000000u
#define CC_PENDING 0xfffffffeu /* a thread is turning this root into a seam root */
#define CC_SENT 0xffffffffu
// A node's second word is (voxel count << 1) | face flag, the count modulo 2^31; wraps are recorded in the carry table of
// the scalar block (components of >= 2^31 voxels: at most two per slab).
__device__ __forceinline__ void cc_carry_add(b2m_scalars *sc, uint32_t root) {
  for (int i = 0; i < 4; i++) {
    const unsigned old = atomicCAS(&sc->carry_slot[i], 0u, root + 1u);
    if (old == 0u || old == root + 1u) { atomicAdd(&sc->carry_hi[i], 1u); return; }
  }
  atomicOr(&sc->overflow, 32u);
}
__device__ __forceinline__ unsigned long long cc_count(const b2m_scalars *sc, uint32_t root, uint32_t y) {
  unsigned long long c = y >> 1;
#pragma unroll
  for (int i = 0; i < 4; i++)
    if (__ldcg(&sc->carry_slot[i]) == root + 1u) c += (unsigned long long)__ldcg(&sc->carry_hi[i]) << 31;
  return c;
}

// Union-find nodes.  A run's slot id is  word*16 + (ordinal of the run inside its word)  - ordered like the
// raster order of the runs' first voxels - but the nodes are STORED with every word's first run in a dense
// array (8 bytes per word, neighbouring words share sectors) and the rare further runs behind it: in smooth
// volumes nearly every access then lands in the dense part instead of pulling one 32-byte sector per 8-byte node.
struct cc_nodes {
  uint2 *p;
  uint32_t nw;  // words of the labelled region
  __device__ __forceinline__ uint2 &operator[](uint32_t slot) const {
    const uint32_t j = slot & 15u, wd = slot >> 4;
    return p[j ? nw + wd * 15u + (j - 1u) : wd];
  }
};
// slot of the run of word `word` (bits wv) that starts at bit s
__device__ __forceinline__ uint32_t run_slot(uint32_t word, uint32_t wv, int s) {
  const uint32_t starts = wv & ~(wv << 1);
  return word * 16u + (uint32_t)__popc(starts & ((1u << s) - 1u));
}

__device__ __forceinline__ uint32_t ld_parent(const cc_nodes &nodes, uint32_t a) {
  return __ldcg(reinterpret_cast<const unsigned int *>(&nodes[a].x));
}

__device__ __forceinline__ uint32_t uf_find(const cc_nodes &nodes, uint32_t a) {
  uint32_t p = ld_parent(nodes, a);
  while (p != a) {
    uint32_t gp = ld_parent(nodes, p);
    if (gp != p) atomicMin(&nodes[a].x, gp);  // path splitting; monotone, so races are benign
    a = p;
    p = gp;
  }
  return a;
}

__device__ __forceinline__ void uf_union(const cc_nodes &nodes, uint32_t a, uint32_t b) {
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
// A CTA labels a tile of CT_W x CT_Y x CT_Z bit words (256 x 8 x 4 voxels; measured: 8x8x8-word tiles label in 3.3 ms
// against 2.8 ms - the barrier between the union and the statistics phase waits for the slowest of 16 warps instead of 8 -
// and the per-tile border pass does not get slower with twice as many z faces: 1.5 vs 1.6 ms) entirely in shared memory
// (same run slots, same union-by-minimum rule, ~30-cycle shared-memory hops instead of L2 round
// trips), then publishes one global node per run: parent = the tile-local root's GLOBAL slot, and on
// the local roots the voxel count and face flag of the whole local component.  Only neighbour pairs
// that straddle two tiles are left for the global union-find (k_cc_border), which therefore works
// on a forest of a few tile-roots per tile instead of one node per run.
#ifndef CT_W
#define CT_W 8
#endif
#ifndef CT_Y
#define CT_Y 8
#endif
#ifndef CT_Z
#define CT_Z 4
#endif
#define CT_WORDS (CT_W * CT_Y * CT_Z)
static_assert(CT_WORDS == 256, "the tile kernels assume 256 threads (eight warps)");

// shared-memory entry of a run slot: parent local slot << 16 | face flag << 15 | voxel count (<= 16384).
// A local slot is  word * 16 + (start bit >> 1);  its entry lives at PIDX(slot) = run-major: in smooth volumes nearly
// every word holds one run with a low start bit, and word-major storage put the 32 lanes of a warp on two banks
// (16-way conflicts on every access: 76 % of the kernel's shared-memory wavefronts in ncu).
#define PIDX(slot) ((((slot)&15u) * CT_WORDS) + ((slot) >> 4))
__device__ __forceinline__ uint32_t lfind(volatile uint32_t *par, uint32_t a) {
  uint32_t p = par[PIDX(a)] >> 16;
  while (p != a) {
    a = p;
    p = par[PIDX(a)] >> 16;
  }
  return a;
}
// find with path halving; only valid while the count bits are still zero (phase A)
__device__ __forceinline__ uint32_t lfind_compress(uint32_t *par, uint32_t a) {
  uint32_t p = ((volatile uint32_t *)par)[PIDX(a)] >> 16;
  while (p != a) {
    const uint32_t gp = ((volatile uint32_t *)par)[PIDX(p)] >> 16;
    if (gp != p) atomicMin(&par[PIDX(a)], gp << 16);
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
    uint32_t old = atomicMin(&par[PIDX(a)], b << 16) >> 16;  // counts are still zero in this phase
    if (old == a) return;
    a = old;
  }
}

// Links that are implied by others are skipped (a solid region would otherwise pay nine unions per word, six of them
// redundant):
//   * word-diagonal link of a wide row (run starts at bit 0 / ends at bit 31 and the neighbour row's previous / next word
//     ends / starts with a set bit): implied when the neighbour row's centre word has bit 0 / bit 31 set - the run links
//     to the centre word's run at that bit, and that run continues into the diagonal word (in-row link);
//   * run of a plane-diagonal row (dy = +-1, dz = -1): implied when, at some voxel where both runs are set, one of the
//     two intermediate face neighbours (dy, 0) or (0, -1) is set too - both runs are face-linked to that voxel's run.
// The implying links are made by this or another thread of this kernel or of its sibling (tile-local / tile-border):
// `fetch` returns 0 for words the calling kernel does not handle, which keeps the test conservative - a link is only
// dropped when the words that imply it are visible to the same kernel.
template <int CONN, class Fetch, class Link>
__device__ __forceinline__ void cc_visit_neighbours(uint32_t rm, int s, int e, Fetch fetch, Link link) {
  auto row = [&](int dy, int dz, bool wide, uint32_t covered) {
    const uint32_t nw = fetch(0, dy, dz);
    uint32_t m = rm;
    if (wide) m |= (rm << 1) | (rm >> 1);
    uint32_t t = nw & m;
    if (nw == 0xffffffffu) {  // a solid word is one run (the common neighbour inside a thick region): no bit scans
      if (!(rm & covered)) link(0, dy, dz, 0, nw);
      t = 0u;
    }
    while (t) {
      const int b = __ffs(t) - 1;
      const int st = run_start(nw, b);
      const int en = run_end(nw, st);
      const uint32_t rb = bits_range(st, en);
      if (!(rb & rm & covered)) link(0, dy, dz, st, nw);
      t &= ~rb;
    }
    if (wide) {
      if (s == 0 && !(nw & 1u)) {
        const uint32_t pw = fetch(-1, dy, dz);
        if (pw >> 31) link(-1, dy, dz, run_start(pw, 31), pw);
      }
      if (e == 31 && !(nw >> 31)) {
        const uint32_t nx = fetch(1, dy, dz);
        if (nx & 1u) link(1, dy, dz, 0, nx);
      }
    }
  };
  if (s == 0) {  // the run continues from the previous word of this row
    const uint32_t pw = fetch(-1, 0, 0);
    if (pw >> 31) link(-1, 0, 0, run_start(pw, 31), pw);
  }
  const bool wide = CONN >= 18;
  row(-1, 0, wide, 0u);
  row(0, -1, wide, 0u);
  if (wide) {
    const uint32_t below = fetch(0, 0, -1);
    row(-1, -1, false, fetch(0, -1, 0) | below);
    row(1, -1, false, fetch(0, 1, 0) | below);
  }
}

// The tile roots (the only nodes that carry counts and can become roots of the global forest) are also appended
// to a compact list, so that the passes after the border unions (k_cc_flatten / k_cc_best) visit a few hundred
// thousand entries instead of scanning every bit word of the volume.  list[0] = count; when it exceeds the
// capacity the full-scan variants of those passes run instead (both are launched, the wrong one exits at once).
#define CT_LIST 1024
// Work compaction: in a smooth volume about half of the words of a tile are empty for either polarity, scattered in
// groups of two or three along x, so with one thread per word every warp carried ~50 % idle lanes through the union,
// statistics and publish phases (ncu r1: 16-19 active threads per instruction).  The non-empty words of the tile are
// therefore compacted into a list first (ballot + prefix over the eight warps) and thread i works on list item i: the
// lanes of the first warps are all busy and the remaining warps skip the phases altogether.
__device__ __forceinline__ unsigned cc_compact_items(bool has, unsigned short *items, unsigned *s_cnt /* [CT_WORDS/32 + 1] */) {
  const unsigned t = threadIdx.x, lane = t & 31u, warp = t >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, has);
  if (lane == 0) s_cnt[warp] = __popc(m);
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (unsigned w = 0; w < CT_WORDS / 32; w++) {
    const unsigned c = s_cnt[w];
    if (w < warp) base += c;
    tot += c;
  }
  if (has) items[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)t;
  __syncthreads();
  return tot;
}

#ifndef CL_MINB
#define CL_MINB 0
#endif
#if CL_MINB
#define CL_BOUNDS __launch_bounds__(CT_WORDS, CL_MINB)
#else
#define CL_BOUNDS __launch_bounds__(CT_WORDS)
#endif
template <int CONN>
__global__ void CL_BOUNDS k_cc_local(const uint32_t *__restrict__ bits, cc_geom g, cc_nodes nodes,
                                                       uint32_t *__restrict__ rlist, unsigned rcap) {
  __shared__ uint32_t sb[CT_WORDS];
  // the same words with a one-word border of zeros below / before / around (z-1, y-1 .. y+1, x-1 .. x+1): a neighbour
  // fetch is one load at a constant offset, without bounds tests (words of other tiles read as empty: k_cc_border's job)
  __shared__ uint32_t sbp[(CT_Z + 1) * (CT_Y + 2) * (CT_W + 2)];
  __shared__ uint32_t par[CT_WORDS * 16];
  __shared__ uint32_t s_list[CT_LIST];
  __shared__ unsigned short s_items[CT_WORDS];
  __shared__ unsigned s_cnt[CT_WORDS / 32 + 1];
  __shared__ unsigned s_n, s_base;
  const int t0 = threadIdx.x;
  if (t0 == 0) s_n = 0;
  const int tx0 = blockIdx.x * CT_W, ty0 = blockIdx.y * CT_Y, tz0 = blockIdx.z * CT_Z;
  unsigned n_items;
  {
    const int lx = t0 % CT_W, ly = (t0 / CT_W) % CT_Y, lz = t0 / (CT_W * CT_Y);
    const int xw = tx0 + lx, y = ty0 + ly, z = tz0 + lz;
    const bool valid = xw < g.w && y < g.ny && z < g.nz;
    const uint32_t wv0 = valid ? __ldg(bits + ((long long)z * g.ny + y) * g.w + xw) : 0u;
    sb[t0] = wv0;
    for (int i = t0; i < (CT_Z + 1) * (CT_Y + 2) * (CT_W + 2); i += CT_WORDS) {
      const int px = i % (CT_W + 2), py = (i / (CT_W + 2)) % (CT_Y + 2), pz = i / ((CT_W + 2) * (CT_Y + 2));
      if (px == 0 || px == CT_W + 1 || py == 0 || py == CT_Y + 1 || pz == 0) sbp[i] = 0u;
    }
    sbp[((lz + 1) * (CT_Y + 2) + (ly + 1)) * (CT_W + 2) + (lx + 1)] = wv0;
    uint32_t starts = wv0 & ~(wv0 << 1);
    while (starts) {
      const int s = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint32_t slot = (uint32_t)t0 * 16u + (uint32_t)(s >> 1);
      par[PIDX(slot)] = slot << 16;
    }
    // uniform tiles: nothing to label in an empty one; a solid one (every word all ones, all inside the volume) is one
    // component whose root is its first word - no unions, no statistics pass
    const bool rowface0 = (y == 0) || (y == g.ny - 1) || (z == 0 && g.zface_lo) || (z == g.nz - 1 && g.zface_hi);
    const int any = __syncthreads_or(wv0 != 0u);
    if (!any) return;
    if (__syncthreads_and(wv0 == 0xffffffffu)) {
      const int faced = __syncthreads_or(rowface0 || xw == 0 || xw * 32 + 31 == g.nx - 1);
      const uint32_t gparent = (uint32_t)(((long long)tz0 * g.ny + ty0) * g.w + tx0) * 16u;
      const uint32_t stat = t0 == 0 ? (((uint32_t)CT_WORDS * 32u) << 1) | (faced ? 1u : 0u) : 0u;
      nodes[(uint32_t)(((long long)z * g.ny + y) * g.w + xw) * 16u] = make_uint2(gparent, stat);
      if (t0 == 0) {
        const unsigned gi = atomicAdd(&rlist[0], 1u);
        if (gi < rcap) rlist[1 + gi] = gparent;
      }
      return;
    }
    // (the list build below synchronises the CTA: sb / par are complete before anybody reads a neighbour's)
    n_items = cc_compact_items(wv0 != 0u, s_items, s_cnt);
  }
  if ((unsigned)(t0 & ~31) >= n_items) {  // this warp has no item: it only takes part in the barriers
    __syncthreads();
    __syncthreads();
  } else {
  const bool act = (unsigned)t0 < n_items;
  const int t = act ? (int)s_items[t0] : 0;  // the word this thread works on
  const int lx = t % CT_W, ly = (t / CT_W) % CT_Y, lz = t / (CT_W * CT_Y);
  const int xw = tx0 + lx, y = ty0 + ly, z = tz0 + lz;
  const long long word = ((long long)z * g.ny + y) * g.w + xw;
  const uint32_t wv = act ? sb[t] : 0u;
  // phase A: unions inside the tile (one thread per non-empty word: handing the neighbour rows of a word to different
  // threads was tried - 2.7x the instructions for the same unions, 4.75 ms instead of 2.8)
  {
    const uint32_t *ctr = sbp + ((lz + 1) * (CT_Y + 2) + (ly + 1)) * (CT_W + 2) + (lx + 1);
    auto fetch = [&](int dx, int dy, int dz) -> uint32_t { return ctr[(dz * (CT_Y + 2) + dy) * (CT_W + 2) + dx]; };
    for (uint32_t rest = wv; rest;) {
      const int s = __ffs(rest) - 1;
      const int e = run_end(wv, s);
      const uint32_t rm = bits_range(s, e);
      rest &= ~rm;
      const uint32_t me = (uint32_t)t * 16u + (uint32_t)(s >> 1);
      cc_visit_neighbours<CONN>(rm, s, e, fetch, [&](int dx, int dy, int dz, int st, uint32_t) {
        const int nt = t + (dz * CT_Y + dy) * CT_W + dx;  // a constant offset at every (inlined) call site
        lunion(par, me, (uint32_t)nt * 16u + (uint32_t)(st >> 1));
      });
    }
  }
  __syncthreads();
  // phase B: local roots collect the voxel count and the face flag of their local component.
  // Lanes that reach the same root combine first (one shared-memory atomic per warp and root: 512
  // same-address atomics per tile were the bottleneck of the first version).
  const bool rowface = (y == 0) || (y == g.ny - 1) || (z == 0 && g.zface_lo) || (z == g.nz - 1 && g.zface_hi);
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
      if (r != 0xffffffffu && (unsigned)(__ffs(peers) - 1) == (unsigned)(t0 & 31)) {
        atomicAdd(&par[PIDX(r)], tot);
        if (fl) atomicOr(&par[PIDX(r)], fl);
      }
    }
  }
  __syncthreads();
  // phase C: publish the global nodes
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
    // local slots are position based (start bit >> 1); the published ids use the run's ordinal in its word
    const uint32_t rwv = sb[rt];
    const uint32_t gparent = (uint32_t)rword * 16u + (uint32_t)__popc((rwv & ~(rwv << 1)) & ((1u << (2u * (r & 15u))) - 1u));
    uint32_t stat = 0;
    if (r == me) {
      const uint32_t pe = par[PIDX(me)];
      stat = ((pe & 0x7fffu) << 1) | ((pe >> 15) & 1u);
      const unsigned li = atomicAdd(&s_n, 1u);
      if (li < CT_LIST) s_list[li] = gparent;
      else {  // more tile roots than the staging list holds (noise): straight to the global list
        const unsigned gi = atomicAdd(&rlist[0], 1u);
        if (gi < rcap) rlist[1 + gi] = gparent;
      }
    }
    nodes[run_slot((uint32_t)word, wv, s)] = make_uint2(gparent, stat);
  }
  }
  __syncthreads();
  const unsigned nl = min(s_n, (unsigned)CT_LIST);
  if (t0 == 0 && nl) s_base = atomicAdd(&rlist[0], nl);
  __syncthreads();
  for (unsigned i = t0; i < nl; i += CT_WORDS)
    if (s_base + i < rcap) rlist[1 + s_base + i] = s_list[i];
}

// neighbour pairs that straddle two tiles: global lock-free unions between (mostly) tile roots
// One CTA per tile.  Only non-empty words on a tile face can have a backward neighbour in another tile: they are
// compacted into a list (as in k_cc_local) and thread i takes item i.  Along the face of two big components every word
// asks for the same (tile root, tile root) pair: lanes of a warp that hold the same pair send one request
// (__match_any_sync), and a small shared-memory cache of recently requested pairs drops the repeats across the
// warps of the tile (whoever put the pair there completes the union inside this kernel).
#define CB_CACHE 128
#ifndef CB_MINB
#define CB_MINB (2048 / CT_WORDS)  /* measured: 1.07 ms at 8 CTAs per SM against 1.25 ms at 6 (the kernel waits on L2) */
#endif
template <int CONN>
__global__ void __launch_bounds__(CT_WORDS, CB_MINB) k_cc_border(const uint32_t *__restrict__ bits, cc_geom g, cc_nodes nodes) {
  __shared__ unsigned long long cache[CB_CACHE];
  __shared__ uint32_t sb[CT_WORDS];
  __shared__ unsigned short s_items[CT_WORDS];
  __shared__ unsigned s_cnt[CT_WORDS / 32 + 1];
  const int t0 = threadIdx.x;
  if (t0 < CB_CACHE) cache[t0] = 0ull;
  unsigned n_items;
  {
    const int lx = t0 % CT_W, ly = (t0 / CT_W) % CT_Y, lz = t0 / (CT_W * CT_Y);
    const int xw = blockIdx.x * CT_W + lx, y = blockIdx.y * CT_Y + ly, z = blockIdx.z * CT_Z + lz;
    const bool face = lx == 0 || lx == CT_W - 1 || ly == 0 || ly == CT_Y - 1 || lz == 0;
    const bool valid = face && xw < g.w && y < g.ny && z < g.nz;
    const uint32_t wv0 = valid ? __ldg(bits + ((long long)z * g.ny + y) * g.w + xw) : 0u;
    sb[t0] = wv0;
    n_items = cc_compact_items(wv0 != 0u, s_items, s_cnt);  // synchronises: cache / sb / list are ready
  }
  if ((unsigned)t0 >= n_items) return;
  const int t = (int)s_items[t0];
  const int lx = t % CT_W, ly = (t / CT_W) % CT_Y, lz = t / (CT_W * CT_Y);
  const int xw = blockIdx.x * CT_W + lx, y = blockIdx.y * CT_Y + ly, z = blockIdx.z * CT_Z + lz;
  const long long word = ((long long)z * g.ny + y) * g.w + xw;
  const uint32_t wv = sb[t];
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
    // unions are made between TILE ROOTS (a run node's parent is its tile root and never changes here); lanes of
    // the warp that want the same pair - the common case along the face of two big components - send one
    const uint32_t ra = nodes[run_slot((uint32_t)word, wv, s)].x;
    cc_visit_neighbours<CONN>(rm, s, e, fetch, [&](int dx, int dy, int dz, int st, uint32_t nwv) {
      const long long nword = ((long long)(z + dz) * g.ny + (y + dy)) * g.w + (xw + dx);
      const uint32_t rb = nodes[run_slot((uint32_t)nword, nwv, st)].x;
      const unsigned long long key = ((unsigned long long)ra << 32) | rb;
      const unsigned peers = __match_any_sync(__activemask(), key);
      if ((unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u)) {
        const unsigned h = ((ra * 0x9E3779B1u) ^ (rb * 0x85EBCA77u)) >> 25;  // 7 bits
        if (atomicExch(&cache[h], key) != key) uf_union(nodes, ra, rb);
      }
    });
  }
}

__device__ __forceinline__ void flush_stats(const cc_nodes &nodes, b2m_scalars *sc, uint32_t root, uint32_t cnt, uint32_t flag) {
  // warp-aggregated: lanes holding the same root combine before touching memory
  unsigned peers = __match_any_sync(__activemask(), root);
  uint32_t tot = __reduce_add_sync(peers, cnt);
  uint32_t fl = __reduce_or_sync(peers, flag);
  if ((unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u) && root != 0xffffffffu) {
    if (tot) {
      const uint32_t old = atomicAdd(&nodes[root].y, tot << 1);
      if ((old >> 1) + tot >= 0x80000000u) cc_carry_add(sc, root);  // the 31-bit count wrapped
    }
    if (fl) atomicOr(&nodes[root].y, 1u);
  }
}

// every tile root that lost its root status in k_cc_border hands the count / face flag of its local
// component to its final root and is pointed straight at it (runs then reach the final root in two hops).
__global__ void __launch_bounds__(256) k_cc_flatten(const uint32_t *__restrict__ bits, cc_geom g, cc_nodes nodes,
                                                    const uint32_t *__restrict__ rlist, unsigned rcap, b2m_scalars *sc) {
  if (rlist[0] <= rcap) return;  // the list variant does the work
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = (g.nwords + 31) / 32 * 32;  // whole warps iterate together
  for (long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x; word < nround; word += stride) {
  uint32_t wv = word < g.nwords ? __ldg(bits + word) : 0u;
  uint32_t starts = wv & ~(wv << 1);
  // all lanes iterate together until every lane is out of runs (keeps the match/reduce converged)
  while (__any_sync(0xffffffffu, starts != 0)) {
    uint32_t root = 0xffffffffu, cnt = 0, flag = 0;
    if (starts) {
      const int s = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint32_t slot = run_slot((uint32_t)word, wv, s);
      const uint2 nd = nodes[slot];
      if ((nd.y >> 1) && nd.x != slot) {  // a tile root (it owns a count) that is no longer a root
        root = uf_find(nodes, slot);
        atomicMin(&nodes[slot].x, root);
        cnt = nd.y >> 1;
        flag = nd.y & 1u;
      }
    }
    if (__any_sync(0xffffffffu, root != 0xffffffffu)) flush_stats(nodes, sc, root, cnt, flag);
  }
  }
}

// list variants: one thread per tile root (k_cc_local's list); they do nothing when the list overflowed
__global__ void __launch_bounds__(256) k_cc_flatten_list(const uint32_t *__restrict__ rlist, unsigned rcap, cc_nodes nodes, b2m_scalars *sc) {
  const unsigned n = rlist[0];
  if (n > rcap) return;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t root = 0xffffffffu, cnt = 0, flag = 0;
  if (i < n) {
    const uint32_t slot = rlist[1 + i];
    const uint2 nd = nodes[slot];
    if (nd.x != slot) {
      root = uf_find(nodes, slot);
      atomicMin(&nodes[slot].x, root);
      cnt = nd.y >> 1;
      flag = nd.y & 1u;
    }
  }
  if (__any_sync(0xffffffffu, root != 0xffffffffu)) flush_stats(nodes, sc, root, cnt, flag);
}
// key of a component = (voxels << 31) | (2^31-1 - root slot): the largest wins, among equals the smallest slot
__device__ __forceinline__ unsigned long long cc_best_key(const b2m_scalars *sc, uint32_t slot, uint32_t y) {
  return (cc_count(sc, slot, y) << 31) | (unsigned long long)(0x7fffffffu - slot);
}
__global__ void __launch_bounds__(256) k_cc_best_list(const uint32_t *__restrict__ rlist, unsigned rcap, cc_nodes nodes,
                                                      unsigned long long *best, unsigned int *nroots, const b2m_scalars *sc) {
  const unsigned n = rlist[0];
  if (n > rcap) return;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long key = 0;
  unsigned int cnt = 0;
  if (i < n) {
    const uint32_t slot = rlist[1 + i];
    const uint2 nd = nodes[slot];
    if (nd.x == slot) {
      cnt = 1;
      key = cc_best_key(sc, slot, nd.y);
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

// final root of a run after k_cc_flatten: run -> tile root -> final root
__device__ __forceinline__ uint32_t cc_final_root(const cc_nodes &nodes, uint32_t slot) {
  uint32_t p = nodes[slot].x;
  while (p != slot) {
    if (p & CC_TAG) return p;  // a seam root: its ticket, CC_PENDING or CC_SENT
    slot = p;
    p = nodes[slot].x;
  }
  return slot;
}

// number of components and the largest one: key = (size << 31) | (2^31-1 - rootslot), so that among equal
// sizes the smallest slot (earliest first voxel in raster order) wins, as src/bwlabel.c:462-466.
__global__ void __launch_bounds__(256) k_cc_best(const uint32_t *__restrict__ bits, long long nwords,
                                                 cc_nodes nodes, unsigned long long *best,
                                                 unsigned int *nroots, const uint32_t *__restrict__ rlist, unsigned rcap,
                                                 const b2m_scalars *sc) {
  if (rlist[0] <= rcap) return;  // the list variant does the work
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long key = 0;
  unsigned int cnt = 0;
  for (long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x; word < nwords; word += stride) {
  uint32_t w = __ldg(bits + word);
  uint32_t starts = w & ~(w << 1);
  while (starts) {
    int s = __ffs(starts) - 1;
    starts &= starts - 1;
    uint32_t slot = run_slot((uint32_t)word, w, s);
    uint2 nd = nodes[slot];
    if (nd.x == slot) {
      cnt++;
      unsigned long long k = cc_best_key(sc, slot, nd.y);
      key = k > key ? k : key;
    }
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
// Slabs: the decision is global and made on the host: sel >= 0 is the winning root slot of THIS rank,
// sel == -1 no local winner (runs of seam roots marked CC_SENT are still selected); sel == -2 reads
// *best (single volume).  nroots_ovr >= 0 replaces *nroots by the global component count.
__global__ void __launch_bounds__(256) k_cc_select(const uint32_t *__restrict__ bits, long long nwords,
                                                   cc_nodes nodes, int mode,
                                                   const unsigned long long *__restrict__ best,
                                                   const unsigned int *__restrict__ nroots,
                                                   const uint32_t *__restrict__ other, uint32_t *__restrict__ out,
                                                   long long sel, long long nroots_ovr) {
  long long word = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (word >= nwords) return;
  uint32_t wv = __ldg(bits + word);
  uint32_t res = 0;
  if (mode == 0) {
    uint32_t bestslot;
    bool have;
    if (sel == -2) { bestslot = 0x7fffffffu - (uint32_t)(*best & 0x7fffffffull); have = *best != 0ull; }
    else { bestslot = sel >= 0 ? (uint32_t)sel : 0xfffffffeu; have = true; }
    uint32_t rest = wv;
    while (rest && have) {
      int s = __ffs(rest) - 1;
      int e = run_end(wv, s);
      uint32_t rm = bits_range(s, e);
      rest &= ~rm;
      const uint32_t r = cc_final_root(nodes, run_slot((uint32_t)word, wv, s));
      if (r == bestslot || r == CC_SENT) res |= rm;
    }
  } else {
    res = __ldg(other + word);
    if ((nroots_ovr >= 0 ? (unsigned long long)nroots_ovr : (unsigned long long)*nroots) > 1ull) {
      uint32_t rest = wv;
      while (rest) {
        int s = __ffs(rest) - 1;
        int e = run_end(wv, s);
        uint32_t rm = bits_range(s, e);
        rest &= ~rm;
        uint32_t root = cc_final_root(nodes, run_slot((uint32_t)word, wv, s));
        if (!(nodes[root].y & 1u)) res |= rm;
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
// Slabs: the arrays are indexed with GLOBAL z (pre-offset EXT buffers), g.nz is the global NZ and the
// kernel writes the own planes [zbeg, zend) only.
// The same pass folds the composition rules (src/meshify.c:332-365) into the marching-cubes inside bits, at
// word level: mb (in: the MC comparison on S; out: the inside bit of the COMPOSED volume) becomes
//   Lewiner  ((mb | fill) & keep | ~keep & [mn inside]) & (~face | [edge_max inside])
//   classic  ((mb & ~fill) & keep | ~keep & [mn inside]) | (face & [edge_max inside])
// because max(v, iso) is never below iso, non-kept voxels hold mn, and the inside test of min(edge_max, v) is
// the AND (Lewiner: v - iso > -eps) / OR (classic: v < iso) of the tests of its two arguments.
struct ibits_params {
  uint32_t *mb;
  uint32_t cM, cE;  // all-ones / zero: inside test of mn, of edge_max
  int classic;
};
/* measured on G1024: 0.338 ms at 8 CTAs per SM (32 registers), 0.345 unconstrained, 0.377 at 1 */
#ifndef DB_MINB
#define DB_MINB 8
#endif
__global__ void __launch_bounds__(256, DB_MINB) k_dilate_bbox(const uint32_t *__restrict__ largest,
                                                     const uint32_t *__restrict__ bright_src, cc_geom g, int zbeg, int zend,
                                                     uint32_t *__restrict__ keep, int *__restrict__ lohi, ibits_params ip) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (y, xw) flattened
  const long long ncol = (long long)g.ny * g.w;
  int lo0 = INT_MAX, lo1 = INT_MAX, lo2 = INT_MAX, hi0 = -1, hi1 = -1, hi2 = -1;
  if (col < ncol) {
    const int y = (int)(col / g.w), xw = (int)(col - (long long)y * g.w);
    const int z0 = zbeg + blockIdx.y * DIL_ZC, z1 = min(z0 + DIL_ZC, zend);
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
      const uint32_t bw = __ldg(bright_src + word);
      if (ip.mb) {
        const uint32_t m = ip.mb[word];
        uint32_t t = ip.classic ? (m & ~bw) : (m | bw);
        t = (t & k) | (~k & ip.cM);
        uint32_t fm = 0xffffffffu;
        if (!(y == 0 || y == g.ny - 1 || z == 0 || z == g.nz - 1)) {
          fm = xw == 0 ? 1u : 0u;
          const int lastb = g.nx - 1 - xw * 32;
          if (lastb >= 0 && lastb < 32) fm |= 1u << lastb;
        }
        t = ip.classic ? (t | (fm & ip.cE)) : (t & (~fm | ip.cE));
        const int nb = g.nx - xw * 32;
        if (nb < 32) t &= (1u << nb) - 1u;
        ip.mb[word] = t;
      }
      const uint32_t bb = bw & k;
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

static unsigned cc_list_cap(const cc_geom &cg) {
  const char *e = getenv("B2M_CC_LIST_CAP");  // tests: a tiny capacity forces the full-scan fallback
  if (e) return (unsigned)atoi(e);
  return (unsigned)(cg.nwords / 4 + 4096);
}

// labelling + number of components (*nroots) + largest component (*best, optional)
static int cc_label(b2m_ctx *ctx, const uint32_t *bits, const cc_geom &cg, cc_nodes nodes, int conn, unsigned long long *best,
                    unsigned int *nroots, b2m_scalars *d_sc) {
  unsigned blocks = b2m_cdiv(cg.nwords, 256);
  dim3 tiles(b2m_cdiv(cg.w, CT_W), b2m_cdiv(cg.ny, CT_Y), b2m_cdiv(cg.nz, CT_Z));
  const unsigned rcap = cc_list_cap(cg);
  B2M_TRY(b2m_reserve(ctx, BUF_CCLIST, ((size_t)rcap + 1) * 4));
  uint32_t *rlist = b2m_ptr<uint32_t>(ctx, BUF_CCLIST);
  CU_TRY(cudaMemsetAsync(rlist, 0, 4, ctx->stream));
  CU_TRY(cudaMemsetAsync(d_sc->carry_slot, 0, 32, ctx->stream));  // carry_slot[4] + carry_hi[4]
  if (conn >= 18) {
    KT_LAUNCH(ctx, "cc_local", k_cc_local<18><<<tiles, CT_WORDS, 0, ctx->stream>>>(bits, cg, nodes, rlist, rcap));
    KT_LAUNCH(ctx, "cc_border", k_cc_border<18><<<tiles, CT_WORDS, 0, ctx->stream>>>(bits, cg, nodes));
  } else {
    KT_LAUNCH(ctx, "cc_local", k_cc_local<6><<<tiles, CT_WORDS, 0, ctx->stream>>>(bits, cg, nodes, rlist, rcap));
    KT_LAUNCH(ctx, "cc_border", k_cc_border<6><<<tiles, CT_WORDS, 0, ctx->stream>>>(bits, cg, nodes));
  }
  const unsigned lblocks = rcap ? b2m_cdiv(rcap, 256) : 1u;
  KT_LAUNCH(ctx, "cc_flatten", k_cc_flatten_list<<<lblocks, 256, 0, ctx->stream>>>(rlist, rcap, nodes, d_sc));
  const unsigned sblocks = min(blocks, (unsigned)ctx->sm_count * 8u);  // grid-stride: these exit at once in the common case
  KT_LAUNCH(ctx, "cc_flatten", k_cc_flatten<<<sblocks, 256, 0, ctx->stream>>>(bits, cg, nodes, rlist, rcap, d_sc));
  KT_LAUNCH(ctx, "cc_best", k_cc_best_list<<<lblocks, 256, 0, ctx->stream>>>(rlist, rcap, nodes, best, nroots, d_sc));
  KT_LAUNCH(ctx, "cc_best", k_cc_best<<<sblocks, 256, 0, ctx->stream>>>(bits, cg.nwords, nodes, best, nroots, rlist, rcap, d_sc));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

// ================================================================================================
// Slabs: merging the components of neighbouring z-slabs.
// Every rank labels its own planes with local run slots.  The local roots that reach a seam plane
// ("seam roots") get compact ids, the upper rank of every seam lists the adjacent (lower root, upper root)
// pairs, pairs and per-root (slot, size, face flag) entries of all ranks are all-gathered, and every rank
// resolves the same small union-find.  Global sizes / face flags come back into the local forest through
// the seam roots only; everything else stays slab-local.
//
// FAST path (default): no sort, no host round trip until the resolution is complete.
//   k_seam_tag       every root that reaches a boundary plane takes a ticket (first come) and is TAGGED: its parent
//                    field becomes CC_TAG | ticket, so the runs of its component resolve to the ticket in O(1);
//                    its entry {slot, flag, size} goes into this rank's block
//   k_seam_dense     tickets of the runs of the last own plane -> dense plane, sent to the rank above
//   k_seam_pairs_f   upper rank: (ticket below, own ticket) pairs, de-duplicated per thread and per CTA
//   all-gather       ONE fixed-size block per rank {header, entries[ent_cap], pairs[pair_cap]}; ids are
//                    rank * ent_cap + ticket, so no offsets have to be agreed on
//   k_seamf_*        replicated union-find over all tickets; component size / face flag / first voxel
//                    (min (rank, slot) = earliest voxel in raster order: src/bwlabel.c:462-466's tie-break)
//   one host sync    sizes, component counts, overflow flags; then k_seamf_apply untags the roots (bubbles: + the
//                    global face flag; largest: CC_SENT on the roots of the winning component)
// A rank whose tickets or pairs exceed the block capacity (noise volumes) raises `overflow`; every rank sees it in
// the gathered headers, all untag and take the SLOW path: sorted unique root lists, variable-length all-gathers
// (the round-1 design, five host round trips).  B2M_SEAM_SLOW=1 forces it, B2M_SEAM_ENT_CAP / _PAIR_CAP shrink the
// block (tests).
// ================================================================================================
struct seam_ent {
  uint32_t slot;            // local root slot (0xffffffff: unused ticket)
  uint32_t flag;            // the local component touches a face of the volume
  unsigned long long cnt;   // its voxels
};
struct seam_hdr {
  unsigned int m, npairs, overflow, pad[13];
};
static_assert(sizeof(seam_hdr) == 64 && sizeof(seam_ent) == 16, "seam block layout");
struct seam_blk {  // view of one rank's block
  seam_hdr *hdr;
  seam_ent *ent;
  uint64_t *pairs;
};
__host__ __device__ static inline size_t seam_blk_bytes(unsigned ent_cap, unsigned pair_cap) {
  return sizeof(seam_hdr) + (size_t)ent_cap * sizeof(seam_ent) + (size_t)pair_cap * 8;
}
__host__ __device__ static inline seam_blk seam_blk_at(char *base, int rank, unsigned ent_cap, unsigned pair_cap) {
  char *p = base + (size_t)rank * seam_blk_bytes(ent_cap, pair_cap);
  seam_blk b;
  b.hdr = reinterpret_cast<seam_hdr *>(p);
  b.ent = reinterpret_cast<seam_ent *>(p + sizeof(seam_hdr));
  b.pairs = reinterpret_cast<uint64_t *>(p + sizeof(seam_hdr) + (size_t)ent_cap * sizeof(seam_ent));
  return b;
}

__global__ void __launch_bounds__(256) k_seam_tag(const uint32_t *__restrict__ bits, cc_geom g, cc_nodes nodes, int do_first, int do_last,
                                                  seam_blk blk, unsigned ent_cap, const b2m_scalars *sc) {
  const long long pw = (long long)g.ny * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pw) return;
  const int which = blockIdx.y;
  if (which == 0 ? !do_first : (!do_last || (g.nz == 1 && do_first))) return;
  const long long word = (which == 0 ? 0 : (long long)(g.nz - 1) * pw) + t;
  const uint32_t wv = __ldg(bits + word);
  uint32_t starts = wv & ~(wv << 1);
  while (starts) {
    const int s = __ffs(starts) - 1;
    starts &= starts - 1;
    const uint32_t r = cc_final_root(nodes, run_slot((uint32_t)word, wv, s));
    if (r & CC_TAG) continue;  // tagged already, or another thread is at it
    if (atomicCAS(&nodes[r].x, r, CC_PENDING) != r) continue;
    const unsigned tk = atomicAdd(&blk.hdr->m, 1u);
    if (tk < ent_cap) {
      const uint32_t y = nodes[r].y;
      seam_ent e;
      e.slot = r; e.flag = y & 1u; e.cnt = cc_count(sc, r, y);
      blk.ent[tk] = e;
      __threadfence();
      atomicExch(&nodes[r].x, CC_TAG | tk);
    } else {
      atomicExch(&nodes[r].x, r);
      atomicOr(&blk.hdr->overflow, 1u);
    }
  }
}
// ticket of the (tagged) root of a run; 0xffffffff when the root carries no ticket (capacity overflow)
__device__ __forceinline__ uint32_t seam_ticket(const cc_nodes &nodes, uint32_t slot) {
  const uint32_t r = cc_final_root(nodes, slot);
  return ((r & CC_TAG) && r < CC_PENDING) ? (r & ~CC_TAG) : 0xffffffffu;
}
__global__ void __launch_bounds__(256) k_seam_dense_f(const uint32_t *__restrict__ bits, cc_geom g, cc_nodes nodes, uint32_t *__restrict__ dense) {
  const long long pw = (long long)g.ny * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pw) return;
  const long long word = (long long)(g.nz - 1) * pw + t;
  const uint32_t wv = __ldg(bits + word);
  uint32_t starts = wv & ~(wv << 1);
  while (starts) {
    const int s = __ffs(starts) - 1;
    starts &= starts - 1;
    dense[(size_t)t * 16 + (s >> 1)] = seam_ticket(nodes, run_slot((uint32_t)word, wv, s));
  }
}
// neighbour runs of run [s,e] (mask rm) of word (xw, y) in the plane `below`: face + (18-connectivity) the four edge
// neighbours of that plane; emit(neighbour word x, neighbour row, start bit of the neighbour run)
template <int CONN, class Emit>
__device__ __forceinline__ void seam_visit_below(const uint32_t *__restrict__ below, const cc_geom &g, int xw, int y, uint32_t rm, int s, int e,
                                                 Emit emit) {
  auto row = [&](int dy, bool wide) {
    const int yy = y + dy;
    if (yy < 0 || yy >= g.ny) return;
    const uint32_t *r = below + (size_t)yy * g.w;
    const uint32_t nw = __ldg(r + xw);
    uint32_t msk = rm;
    if (wide) msk |= (rm << 1) | (rm >> 1);
    uint32_t tt = nw & msk;
    while (tt) {
      const int b = __ffs(tt) - 1;
      const int st = run_start(nw, b);
      const int en = run_end(nw, st);
      emit(xw, yy, st);
      tt &= ~bits_range(st, en);
    }
    if (wide) {
      if (s == 0 && xw > 0) {
        const uint32_t pv = __ldg(r + xw - 1);
        if (pv >> 31) emit(xw - 1, yy, run_start(pv, 31));
      }
      if (e == 31 && xw + 1 < g.w) {
        const uint32_t nv = __ldg(r + xw + 1);
        if (nv & 1u) emit(xw + 1, yy, 0);
      }
    }
  };
  row(0, CONN >= 18);
  if (CONN >= 18) { row(-1, false); row(1, false); }
}
#define SEAM_CACHE 256
template <int CONN>
__global__ void __launch_bounds__(256) k_seam_pairs_f(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ below, cc_geom g,
                                                      cc_nodes nodes, const uint32_t *__restrict__ dense_below, unsigned id_me,
                                                      unsigned id_below, seam_blk blk, unsigned pair_cap) {
  __shared__ unsigned long long cache[SEAM_CACHE];
  cache[threadIdx.x] = ~0ull;
  __syncthreads();
  const long long pw = (long long)g.ny * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pw) return;
  const int y = (int)(t / g.w), xw = (int)(t - (long long)y * g.w);
  const uint32_t wv = __ldg(bits + t);
  uint64_t last = ~0ull;
  for (uint32_t rest = wv; rest;) {
    const int s = __ffs(rest) - 1;
    const int e = run_end(wv, s);
    const uint32_t rm = bits_range(s, e);
    rest &= ~rm;
    const uint32_t tk = seam_ticket(nodes, run_slot((uint32_t)t, wv, s));
    if (tk == 0xffffffffu) continue;  // overflow: the slow path takes over
    const uint64_t me = (uint64_t)(id_me + tk);
    seam_visit_below<CONN>(below, g, xw, y, rm, s, e, [&](int nxw, int ny_, int st) {
      const uint32_t tb = __ldg(dense_below + ((size_t)ny_ * g.w + nxw) * 16 + (st >> 1));
      if (tb == 0xffffffffu) return;
      const uint64_t key = ((uint64_t)(id_below + tb) << 32) | me;
      if (key == last) return;
      last = key;
      const unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 56);
      if (atomicExch(&cache[h], key) == key) return;  // this CTA has just listed the pair
      const unsigned pos = atomicAdd(&blk.hdr->npairs, 1u);
      if (pos < pair_cap) blk.pairs[pos] = key; else atomicOr(&blk.hdr->overflow, 2u);
    });
  }
}

__device__ __forceinline__ uint32_t uf32_find(uint32_t *par, uint32_t a) {
  uint32_t p = __ldcg(par + a);
  while (p != a) {
    const uint32_t gp = __ldcg(par + p);
    if (gp != p) atomicMin(par + a, gp);
    a = p;
    p = gp;
  }
  return a;
}
__device__ __forceinline__ void uf32_union(uint32_t *par, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf32_find(par, a);
    b = uf32_find(par, b);
    if (a == b) return;
    if (a < b) { const uint32_t t = a; a = b; b = t; }
    const uint32_t old = atomicMin(par + a, b);
    if (old == a) return;
    a = old;
  }
}
// replicated resolution over the gathered blocks (ids: rank * ent_cap + ticket)
struct seamf_arrays {
  char *blocks;                 // world blocks
  uint32_t *par;                // [world * ent_cap]
  unsigned long long *gcnt;     // component size, at the root id
  unsigned long long *gprio;    // min (rank << 32 | slot) of the component, at the root id
  uint32_t *gflag;              // component touches a face
  int world;
  unsigned ent_cap, pair_cap;
};
__global__ void __launch_bounds__(256) k_seamf_init(seamf_arrays a) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)a.world * a.ent_cap) return;
  a.par[i] = i; a.gcnt[i] = 0ull; a.gprio[i] = ~0ull; a.gflag[i] = 0u;
}
__global__ void __launch_bounds__(256) k_seamf_union(seamf_arrays a) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)a.world * a.pair_cap) return;
  const int r = (int)(i / a.pair_cap);
  const unsigned j = i - (unsigned)r * a.pair_cap;
  const seam_blk b = seam_blk_at(a.blocks, r, a.ent_cap, a.pair_cap);
  if (j >= min(b.hdr->npairs, a.pair_cap)) return;
  const uint64_t key = b.pairs[j];
  uf32_union(a.par, (uint32_t)(key >> 32), (uint32_t)(key & 0xffffffffull));
}
__global__ void __launch_bounds__(256) k_seamf_stats(seamf_arrays a) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)a.world * a.ent_cap) return;
  const int r = (int)(i / a.ent_cap);
  const unsigned t = i - (unsigned)r * a.ent_cap;
  const seam_blk b = seam_blk_at(a.blocks, r, a.ent_cap, a.pair_cap);
  if (t >= min(b.hdr->m, a.ent_cap)) return;
  const seam_ent e = b.ent[t];
  const uint32_t root = uf32_find(a.par, i);  // unions are complete (previous kernel): this is the final root
  if (root != i) atomicMin(a.par + i, root);
  atomicAdd(a.gcnt + root, e.cnt);
  if (e.flag) atomicOr(a.gflag + root, 1u);
  atomicMin(a.gprio + root, ((unsigned long long)r << 32) | e.slot);
}
// one CTA: number of seam roots / components among them, the largest component (size, then earliest first voxel),
// and whether any rank overflowed its block
__global__ void __launch_bounds__(1024) k_seamf_best(seamf_arrays a, b2m_scalars *sc) {
  __shared__ unsigned long long s_size, s_prio;
  __shared__ unsigned s_m, s_k, s_over;
  if (threadIdx.x == 0) { s_size = 0ull; s_prio = ~0ull; s_m = 0u; s_k = 0u; s_over = 0u; }
  __syncthreads();
  const unsigned n = (unsigned)a.world * a.ent_cap;
  unsigned long long mx = 0ull;
  unsigned m = 0, k = 0;
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = (int)(i / a.ent_cap);
    const seam_blk b = seam_blk_at(a.blocks, r, a.ent_cap, a.pair_cap);
    if (i - (unsigned)r * a.ent_cap >= min(b.hdr->m, a.ent_cap)) continue;
    m++;
    if (a.par[i] == i) { k++; const unsigned long long c = a.gcnt[i]; mx = c > mx ? c : mx; }
  }
  if (mx) atomicMax(&s_size, mx);
  if (m) atomicAdd(&s_m, m);
  if (k) atomicAdd(&s_k, k);
  if (threadIdx.x < (unsigned)a.world && seam_blk_at(a.blocks, threadIdx.x, a.ent_cap, a.pair_cap).hdr->overflow) atomicOr(&s_over, 1u);
  __syncthreads();
  const unsigned long long best = s_size;
  unsigned long long pr = ~0ull;
  if (best)
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
      const int r = (int)(i / a.ent_cap);
      const seam_blk b = seam_blk_at(a.blocks, r, a.ent_cap, a.pair_cap);
      if (i - (unsigned)r * a.ent_cap >= min(b.hdr->m, a.ent_cap)) continue;
      if (a.par[i] == i && a.gcnt[i] == best) { const unsigned long long p = a.gprio[i]; pr = p < pr ? p : pr; }
    }
  if (pr != ~0ull) atomicMin(&s_prio, pr);
  __syncthreads();
  if (threadIdx.x == 0) {
    sc->seam_n = s_m; sc->seam_roots = s_k; sc->best_seam = best; sc->best_seam_prio = s_prio;
    if (s_over) atomicOr(&sc->overflow, 64u);
  }
}
// mode 2: untag only; mode 1: untag + the component's global face flag; mode 0: untag, CC_SENT on the roots of the
// component whose first voxel is gstar_prio
__global__ void __launch_bounds__(256) k_seamf_apply(seamf_arrays a, int me, cc_nodes nodes, int mode, unsigned long long gstar_prio) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const seam_blk b = seam_blk_at(a.blocks, me, a.ent_cap, a.pair_cap);
  if (i >= min(b.hdr->m, a.ent_cap)) return;
  const uint32_t slot = b.ent[i].slot;
  const uint32_t root = mode == 2 ? 0u : a.par[(unsigned)me * a.ent_cap + i];
  if (mode == 1 && a.gflag[root]) atomicOr(&nodes[slot].y, 1u);
  nodes[slot].x = (mode == 0 && a.gprio[root] == gstar_prio) ? CC_SENT : slot;
}

// ---- slow path kernels (sorted unique lists) ----
__global__ void __launch_bounds__(256) k_seam_collect(const uint32_t *__restrict__ bits, cc_geom g, cc_nodes nodes,
                                                      int do_first, int do_last, uint64_t *__restrict__ list, unsigned cap,
                                                      unsigned int *__restrict__ count, unsigned int *__restrict__ overflow) {
  const long long pw = (long long)g.ny * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pw) return;
  const int which = blockIdx.y;
  if (which == 0 ? !do_first : (!do_last || (g.nz == 1 && do_first))) return;
  const long long word = (which == 0 ? 0 : (long long)(g.nz - 1) * pw) + t;
  const uint32_t wv = __ldg(bits + word);
  uint32_t starts = wv & ~(wv << 1);
  while (starts) {
    const int s = __ffs(starts) - 1;
    starts &= starts - 1;
    const uint32_t root = cc_final_root(nodes, run_slot((uint32_t)word, wv, s));
    const unsigned pos = atomicAdd(count, 1u);
    if (pos < cap) list[pos] = root; else atomicOr(overflow, 4u);
  }
}
__global__ void __launch_bounds__(256) k_seam_unique_flags(const uint64_t *__restrict__ keys, unsigned n, uint32_t *__restrict__ flag) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
template <typename T>
__global__ void __launch_bounds__(256) k_seam_unique_scatter(const uint64_t *__restrict__ keys, unsigned n, const uint32_t *__restrict__ scan,
                                                             const unsigned int *__restrict__ total, T *__restrict__ out) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t me = scan[i], next = i + 1 < n ? scan[i + 1] : *total;
  if (next != me) out[me] = (T)keys[i];
}
__device__ __forceinline__ unsigned seam_lower_bound(const uint32_t *__restrict__ a, unsigned n, uint32_t v) {
  unsigned lo = 0, hi = n;
  while (lo < hi) {
    const unsigned mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// dense[word in plane * 16 + run slot] = index of the run's root in U, for the LAST own plane (sent up)
__global__ void __launch_bounds__(256) k_seam_dense(const uint32_t *__restrict__ bits, cc_geom g, cc_nodes nodes,
                                                    const uint32_t *__restrict__ U, unsigned m, uint32_t *__restrict__ dense) {
  const long long pw = (long long)g.ny * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pw) return;
  const long long word = (long long)(g.nz - 1) * pw + t;
  const uint32_t wv = __ldg(bits + word);
  uint32_t starts = wv & ~(wv << 1);
  while (starts) {
    const int s = __ffs(starts) - 1;
    starts &= starts - 1;
    const uint32_t root = cc_final_root(nodes, run_slot((uint32_t)word, wv, s));
    dense[(size_t)t * 16 + (s >> 1)] = seam_lower_bound(U, m, root);
  }
}
// upper rank of a seam: every run of the first own plane against the runs of the plane below
template <int CONN>
__global__ void __launch_bounds__(256) k_seam_pairs(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ below, cc_geom g,
                                                    cc_nodes nodes, const uint32_t *__restrict__ U, unsigned m,
                                                    unsigned off_me, unsigned off_below, const uint32_t *__restrict__ dense_below,
                                                    int mb, uint64_t *__restrict__ pairs, unsigned cap,
                                                    unsigned int *__restrict__ count, unsigned int *__restrict__ overflow) {
  const long long pw = (long long)g.ny * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pw) return;
  const int y = (int)(t / g.w), xw = (int)(t - (long long)y * g.w);
  const uint32_t wv = __ldg(bits + t);
  uint64_t last = ~0ull;
  for (uint32_t rest = wv; rest;) {
    const int s = __ffs(rest) - 1;
    const int e = run_end(wv, s);
    const uint32_t rm = bits_range(s, e);
    rest &= ~rm;
    const uint32_t root = cc_final_root(nodes, run_slot((uint32_t)t, wv, s));
    const uint64_t me = (uint64_t)(off_me + seam_lower_bound(U, m, root));
    seam_visit_below<CONN>(below, g, xw, y, rm, s, e, [&](int nxw, int ny_, int st) {
      const uint64_t nb = (uint64_t)(off_below + __ldg(dense_below + ((size_t)ny_ * g.w + nxw) * 16 + (st >> 1)));
      const uint64_t key = (nb << mb) | me;
      if (key == last) return;
      last = key;
      const unsigned pos = atomicAdd(count, 1u);
      if (pos < cap) pairs[pos] = key; else atomicOr(overflow, 8u);
    });
  }
}
// per own seam root: {local slot, face flag, voxels}
__global__ void __launch_bounds__(256) k_seam_entries(const uint32_t *__restrict__ U, unsigned m, cc_nodes nodes, const b2m_scalars *sc,
                                                      seam_ent *__restrict__ ent) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t y = nodes[U[i]].y;
  seam_ent e;
  e.slot = U[i]; e.flag = y & 1u; e.cnt = cc_count(sc, U[i], y);
  ent[i] = e;
}
__global__ void __launch_bounds__(256) k_seam_iota(uint32_t *__restrict__ par, unsigned n) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) par[i] = i;
}
__global__ void __launch_bounds__(256) k_seam_union(const uint64_t *__restrict__ pairs, unsigned n, int mb, uint32_t *par) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uf32_union(par, (uint32_t)(pairs[i] >> mb), (uint32_t)(pairs[i] & ((1ull << mb) - 1ull)));
}
__global__ void __launch_bounds__(256) k_seam_flatten(uint32_t *par, unsigned n) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = uf32_find(par, i);
  if (r != i) atomicMin(par + i, r);
}
__global__ void __launch_bounds__(256) k_seam_stats(unsigned n, const uint32_t *__restrict__ par, const seam_ent *__restrict__ ent,
                                                    unsigned long long *__restrict__ gcnt, uint32_t *__restrict__ gflag) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = par[i];
  atomicAdd(gcnt + r, ent[i].cnt);
  if (ent[i].flag) atomicOr(gflag + r, 1u);
}
__global__ void __launch_bounds__(256) k_seam_best(unsigned n, const uint32_t *__restrict__ par, const unsigned long long *__restrict__ gcnt,
                                                   unsigned long long *__restrict__ best, unsigned int *__restrict__ nroots) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || par[i] != i) return;
  atomicAdd(nroots, 1u);
  atomicMax(best, (gcnt[i] << 30) | (unsigned long long)(0x3fffffffu - i));
}
// mode 1: global face flags into the own seam roots; mode 0: mark the own seam roots of component gstar
__global__ void __launch_bounds__(256) k_seam_apply(unsigned m, unsigned off, const uint32_t *__restrict__ U, const uint32_t *__restrict__ par,
                                                    const uint32_t *__restrict__ gflag, cc_nodes nodes, int mode, long long gstar) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t r = par[off + i];
  if (mode == 1) { if (gflag[r]) atomicOr(&nodes[U[i]].y, 1u); }
  else if ((long long)r == gstar) nodes[U[i]].x = CC_SENT;
}

__global__ void k_clear_bits(unsigned int *p, unsigned int bits) { atomicAnd(p, ~bits); }

struct seam_result {
  int fast;                     // which path produced it
  unsigned M, K;                // seam roots over all ranks, components among them
  unsigned long long best_size; // voxels of the largest seam component (0: none)
  int best_rank; unsigned best_slot;  // first voxel of that component: (rank, local slot)
  // fast path
  seamf_arrays fa;
  unsigned long long best_prio;
  // slow path
  long long gstar;              // id of the largest seam component
  unsigned m, off;              // own seam roots and their id offset
  const uint32_t *U, *par, *gflag;
};

static int seam_bits(unsigned long long n) {
  int b = 1;
  while ((1ull << b) < n) b++;
  return b;
}
static unsigned seam_env_cap(const char *name, unsigned dflt) {
  const char *e = getenv(name);
  if (!e) return dflt;
  const long v = atol(e);
  return v >= 1 && v <= (1 << 20) ? (unsigned)v : dflt;
}

// labelled `bits_own` (own planes, geometry cg) -> seam resolution.  `below` = the neighbour's last plane
// (EXT plane 0) when this rank has a lower neighbour.  Returns 1 in *overflowed when some rank's block was too small
// (the roots are untagged again; nothing else has changed).
static int cc_seams_fast(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const uint32_t *bits_own, const uint32_t *below,
                         const cc_geom &cg, cc_nodes nodes, int conn, b2m_scalars *d_sc, seam_result *sr, int *overflowed) {
  const int W = sl.world, me = sl.rank;
  const size_t pw = (size_t)cg.ny * cg.w;
  const unsigned pblocks = b2m_cdiv(pw, 256);
  const unsigned ent_cap = seam_env_cap("B2M_SEAM_ENT_CAP", 8192), pair_cap = seam_env_cap("B2M_SEAM_PAIR_CAP", 32768);
  const size_t bb = seam_blk_bytes(ent_cap, pair_cap);
  const size_t M = (size_t)W * ent_cap;
  memset(sr, 0, sizeof(*sr));
  sr->fast = 1;
  *overflowed = 0;
  size_t o = 0;
  auto carve = [&](size_t bytes) { const size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
  const size_t o_blk = carve(bb * W), o_par = carve(M * 4), o_cnt = carve(M * 8), o_prio = carve(M * 8), o_flag = carve(M * 4);
  const size_t o_ds = carve(pw * 64), o_dr = carve(pw * 64);
  B2M_TRY(b2m_reserve(ctx, BUF_SEAMF, o));
  char *base = b2m_ptr<char>(ctx, BUF_SEAMF);
  seamf_arrays fa;
  fa.blocks = base + o_blk; fa.par = reinterpret_cast<uint32_t *>(base + o_par);
  fa.gcnt = reinterpret_cast<unsigned long long *>(base + o_cnt); fa.gprio = reinterpret_cast<unsigned long long *>(base + o_prio);
  fa.gflag = reinterpret_cast<uint32_t *>(base + o_flag);
  fa.world = W; fa.ent_cap = ent_cap; fa.pair_cap = pair_cap;
  uint32_t *dense_send = reinterpret_cast<uint32_t *>(base + o_ds), *dense_recv = reinterpret_cast<uint32_t *>(base + o_dr);
  const seam_blk mine = seam_blk_at(fa.blocks, me, ent_cap, pair_cap);
  CU_TRY(cudaMemsetAsync(mine.hdr, 0, sizeof(seam_hdr), ctx->stream));
  KT_LAUNCH(ctx, "seam_tag", k_seam_tag<<<dim3(pblocks, 2), 256, 0, ctx->stream>>>(bits_own, cg, nodes, sl.hl, sl.hh, mine, ent_cap, d_sc));
  if (sl.hh) KT_LAUNCH(ctx, "seam_dense", k_seam_dense_f<<<pblocks, 256, 0, ctx->stream>>>(bits_own, cg, nodes, dense_send));  // only run-start slots are written (and read)
  B2M_TRY(b2m_comm_exchange(ctx, comm, dense_send, sl.hh ? pw * 64 : 0, dense_recv, sl.hl ? pw * 64 : 0, nullptr, 0, nullptr, 0));
  if (sl.hl) {
    if (conn >= 18)
      KT_LAUNCH(ctx, "seam_pairs", k_seam_pairs_f<18><<<pblocks, 256, 0, ctx->stream>>>(bits_own, below, cg, nodes, dense_recv, (unsigned)me * ent_cap, (unsigned)(me - 1) * ent_cap, mine, pair_cap));
    else
      KT_LAUNCH(ctx, "seam_pairs", k_seam_pairs_f<6><<<pblocks, 256, 0, ctx->stream>>>(bits_own, below, cg, nodes, dense_recv, (unsigned)me * ent_cap, (unsigned)(me - 1) * ent_cap, mine, pair_cap));
  }
  B2M_TRY(b2m_comm_allgather_inplace(ctx, comm, fa.blocks, bb));
  KT_LAUNCH(ctx, "seam_uf", k_seamf_init<<<b2m_cdiv(M, 256), 256, 0, ctx->stream>>>(fa));
  KT_LAUNCH(ctx, "seam_uf", k_seamf_union<<<b2m_cdiv((size_t)W * pair_cap, 256), 256, 0, ctx->stream>>>(fa));
  KT_LAUNCH(ctx, "seam_stats", k_seamf_stats<<<b2m_cdiv(M, 256), 256, 0, ctx->stream>>>(fa));
  KT_LAUNCH(ctx, "seam_stats", k_seamf_best<<<1, 1024, 0, ctx->stream>>>(fa, d_sc));
  CU_TRY(cudaGetLastError());
  B2M_TRY(b2m_sync_scalars(ctx, comm));  // the one host round trip of the fast path (also carries nroots / best of the local labelling)
  bool over = false;
  for (int r = 0; r < W; r++) over |= (b2m_sc(ctx, comm, r)->overflow & 64u) != 0;  // every rank computed it from the same headers
  sr->fa = fa;
  if (over) {
    KT_LAUNCH(ctx, "seam_apply", k_seamf_apply<<<b2m_cdiv(ent_cap, 256), 256, 0, ctx->stream>>>(fa, me, nodes, 2, 0ull));
    k_clear_bits<<<1, 1, 0, ctx->stream>>>(&d_sc->overflow, 64u);
    *overflowed = 1;
    return B2M_OK;
  }
  const b2m_scalars *h = ctx->h_scalars;
  sr->M = h->seam_n; sr->K = h->seam_roots;
  sr->best_size = h->best_seam; sr->best_prio = h->best_seam_prio;
  sr->best_rank = (int)(h->best_seam_prio >> 32); sr->best_slot = (unsigned)(h->best_seam_prio & 0xffffffffull);
  return B2M_OK;
}
// give the global result back to the local forest (and untag): mode 1 face flags, mode 0 CC_SENT on the winner (win: the
// seam component won the largest-cluster contest), mode 2 nothing
static int cc_seams_apply(b2m_ctx *ctx, const b2m_slab &sl, cc_nodes nodes, const seam_result *sr, int mode, bool win) {
  if (sr->fast) {
    KT_LAUNCH(ctx, "seam_apply", k_seamf_apply<<<b2m_cdiv(sr->fa.ent_cap, 256), 256, 0, ctx->stream>>>(sr->fa, sl.rank, nodes, mode == 0 && !win ? 2 : mode, sr->best_prio));
  } else if (sr->m && (mode == 1 || (mode == 0 && win))) {
    KT_LAUNCH(ctx, "seam_apply", k_seam_apply<<<b2m_cdiv(sr->m, 256), 256, 0, ctx->stream>>>(sr->m, sr->off, sr->U, sr->par, sr->gflag, nodes, mode, sr->gstar));
  }
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

static int cc_seams_slow(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const uint32_t *bits_own, const uint32_t *below,
                         const cc_geom &cg, cc_nodes nodes, int conn, b2m_scalars *d_sc, seam_result *sr) {
  const int W = sl.world, me = sl.rank;
  const size_t pw = (size_t)cg.ny * cg.w;
  const unsigned pblocks = b2m_cdiv(pw, 256);
  memset(sr, 0, sizeof(*sr));
  sr->gstar = -1;
  // 1. roots of the runs on the own boundary planes -> sorted unique list U
  const unsigned cap_runs = (unsigned)(2 * pw * 16);
  B2M_TRY(b2m_reserve(ctx, BUF_SEAM0, (size_t)cap_runs * 8 + (size_t)cap_runs * 4 + 256));
  uint64_t *list = b2m_ptr<uint64_t>(ctx, BUF_SEAM0);
  uint32_t *flag = reinterpret_cast<uint32_t *>(list + cap_runs);
  B2M_TRY(b2m_reserve(ctx, BUF_SEAM1, (size_t)cap_runs * 4 + 2 * pw * 64 + 256));
  uint32_t *U = b2m_ptr<uint32_t>(ctx, BUF_SEAM1);
  uint32_t *dense_send = U + cap_runs, *dense_recv = dense_send + pw * 16;
  CU_TRY(cudaMemsetAsync(&d_sc->seam_n, 0, 4, ctx->stream));
  KT_LAUNCH(ctx, "seam_collect", k_seam_collect<<<dim3(pblocks, 2), 256, 0, ctx->stream>>>(bits_own, cg, nodes, sl.hl, sl.hh, list, cap_runs, &d_sc->seam_n, &d_sc->overflow));
  CU_TRY(cudaGetLastError());
  B2M_TRY(b2m_fetch_scalars(ctx));
  const unsigned nL = ctx->h_scalars->seam_n;
  if (nL > cap_runs) { b2m_set_error("seam: run list overflow"); return B2M_ECUDA; }
  if (nL > 0) {
    B2M_TRY(b2m_sort_u64(ctx, list, nL, 32));
    KT_LAUNCH(ctx, "seam_unique", k_seam_unique_flags<<<b2m_cdiv(nL, 256), 256, 0, ctx->stream>>>(list, nL, flag));
    B2M_TRY(b2m_exclusive_scan_u32(ctx, flag, flag, nL, &d_sc->seam_n));
    KT_LAUNCH(ctx, "seam_unique", k_seam_unique_scatter<uint32_t><<<b2m_cdiv(nL, 256), 256, 0, ctx->stream>>>(list, nL, flag, &d_sc->seam_n, U));
  }
  CU_TRY(cudaGetLastError());
  B2M_TRY(b2m_sync_scalars(ctx, comm));
  unsigned offs[65], M = 0;
  for (int r = 0; r < W; r++) { offs[r] = M; M += b2m_sc(ctx, comm, r)->seam_n; }
  offs[W] = M;
  const unsigned m = offs[me + 1] - offs[me];
  sr->M = M; sr->m = m; sr->off = offs[me]; sr->U = U;
  if (M == 0) return B2M_OK;
  if (M >= (1u << 30)) { b2m_set_error("seam: too many seam components"); return B2M_EARG; }
  const int mb = seam_bits(M);
  // 2. compact ids of the last own plane go up; pairs are listed by the upper rank of each seam
  if (sl.hh) {
    CU_TRY(cudaMemsetAsync(dense_send, 0, pw * 64, ctx->stream));  // only run-start slots are written (and read)
    KT_LAUNCH(ctx, "seam_dense", k_seam_dense<<<pblocks, 256, 0, ctx->stream>>>(bits_own, cg, nodes, U, m, dense_send));
  }
  B2M_TRY(b2m_comm_exchange(ctx, comm, dense_send, sl.hh ? pw * 64 : 0, dense_recv, sl.hl ? pw * 64 : 0, nullptr, 0, nullptr, 0));
  const unsigned cap_pairs = (unsigned)(pw * 96);
  B2M_TRY(b2m_reserve(ctx, BUF_SEAM2, (size_t)cap_pairs * 8 + (size_t)cap_pairs * 4 + 256));
  uint64_t *pairs = b2m_ptr<uint64_t>(ctx, BUF_SEAM2);
  uint32_t *pflag = reinterpret_cast<uint32_t *>(pairs + cap_pairs);
  CU_TRY(cudaMemsetAsync(&d_sc->seam_n, 0, 4, ctx->stream));
  if (sl.hl) {
    if (conn >= 18)
      KT_LAUNCH(ctx, "seam_pairs", k_seam_pairs<18><<<pblocks, 256, 0, ctx->stream>>>(bits_own, below, cg, nodes, U, m, offs[me], offs[me - 1], dense_recv, mb, pairs, cap_pairs, &d_sc->seam_n, &d_sc->overflow));
    else
      KT_LAUNCH(ctx, "seam_pairs", k_seam_pairs<6><<<pblocks, 256, 0, ctx->stream>>>(bits_own, below, cg, nodes, U, m, offs[me], offs[me - 1], dense_recv, mb, pairs, cap_pairs, &d_sc->seam_n, &d_sc->overflow));
  }
  CU_TRY(cudaGetLastError());
  B2M_TRY(b2m_fetch_scalars(ctx));
  const unsigned np_raw = ctx->h_scalars->seam_n;
  if (np_raw > cap_pairs) { b2m_set_error("seam: pair list overflow"); return B2M_ECUDA; }
  // the unique pairs and the own entries go to BUF_SEAM0 (the root list is no longer needed)
  const size_t up_bytes = ((size_t)np_raw * 8 + 255) & ~(size_t)255;
  B2M_TRY(b2m_reserve(ctx, BUF_SEAM0, up_bytes + (size_t)m * sizeof(seam_ent) + 256));
  uint64_t *upairs = b2m_ptr<uint64_t>(ctx, BUF_SEAM0);
  seam_ent *ent_own = reinterpret_cast<seam_ent *>(b2m_ptr<char>(ctx, BUF_SEAM0) + up_bytes);
  if (np_raw > 0) {
    B2M_TRY(b2m_sort_u64(ctx, pairs, np_raw, 2 * mb));
    KT_LAUNCH(ctx, "seam_unique", k_seam_unique_flags<<<b2m_cdiv(np_raw, 256), 256, 0, ctx->stream>>>(pairs, np_raw, pflag));
    B2M_TRY(b2m_exclusive_scan_u32(ctx, pflag, pflag, np_raw, &d_sc->seam_n));
    KT_LAUNCH(ctx, "seam_unique", k_seam_unique_scatter<uint64_t><<<b2m_cdiv(np_raw, 256), 256, 0, ctx->stream>>>(pairs, np_raw, pflag, &d_sc->seam_n, upairs));
  }
  CU_TRY(cudaGetLastError());
  B2M_TRY(b2m_sync_scalars(ctx, comm));
  size_t pbytes[64], ebytes[64];
  unsigned Ptot = 0;
  for (int r = 0; r < W; r++) {
    const unsigned npr = b2m_sc(ctx, comm, r)->seam_n;
    pbytes[r] = (size_t)npr * 8;
    ebytes[r] = (size_t)(offs[r + 1] - offs[r]) * sizeof(seam_ent);
    Ptot += npr;
  }
  // 3. replicated resolution.  BUF_SEAM2 is re-carved: all pairs | all entries | parent | gcnt | gflag
  size_t need = (size_t)Ptot * 8 + (size_t)M * (sizeof(seam_ent) + 4 + 8 + 4) + 2048;
  B2M_TRY(b2m_reserve(ctx, BUF_SEAM2, need));
  char *base = b2m_ptr<char>(ctx, BUF_SEAM2);
  size_t o = 0;
  auto take = [&](size_t bytes) { char *q = base + o; o = (o + bytes + 255) & ~(size_t)255; return q; };
  uint64_t *all_pairs = reinterpret_cast<uint64_t *>(take((size_t)Ptot * 8));
  seam_ent *all_ent = reinterpret_cast<seam_ent *>(take((size_t)M * sizeof(seam_ent)));
  uint32_t *par = reinterpret_cast<uint32_t *>(take((size_t)M * 4));
  unsigned long long *gcnt = reinterpret_cast<unsigned long long *>(take((size_t)M * 8));
  uint32_t *gflag = reinterpret_cast<uint32_t *>(take((size_t)M * 4));
  if (m) KT_LAUNCH(ctx, "seam_entries", k_seam_entries<<<b2m_cdiv(m, 256), 256, 0, ctx->stream>>>(U, m, nodes, d_sc, ent_own));
  B2M_TRY(b2m_comm_allgatherv(ctx, comm, upairs, all_pairs, pbytes));
  B2M_TRY(b2m_comm_allgatherv(ctx, comm, ent_own, all_ent, ebytes));
  const unsigned mblocks = b2m_cdiv(M, 256);
  KT_LAUNCH(ctx, "seam_uf", k_seam_iota<<<mblocks, 256, 0, ctx->stream>>>(par, M));
  if (Ptot) KT_LAUNCH(ctx, "seam_uf", k_seam_union<<<b2m_cdiv(Ptot, 256), 256, 0, ctx->stream>>>(all_pairs, Ptot, mb, par));
  KT_LAUNCH(ctx, "seam_uf", k_seam_flatten<<<mblocks, 256, 0, ctx->stream>>>(par, M));
  CU_TRY(cudaMemsetAsync(gcnt, 0, (size_t)M * 8, ctx->stream));
  CU_TRY(cudaMemsetAsync(gflag, 0, (size_t)M * 4, ctx->stream));
  KT_LAUNCH(ctx, "seam_stats", k_seam_stats<<<mblocks, 256, 0, ctx->stream>>>(M, par, all_ent, gcnt, gflag));
  CU_TRY(cudaMemsetAsync(&d_sc->best_seam, 0, 8, ctx->stream));
  CU_TRY(cudaMemsetAsync(&d_sc->seam_roots, 0, 4, ctx->stream));
  KT_LAUNCH(ctx, "seam_stats", k_seam_best<<<mblocks, 256, 0, ctx->stream>>>(M, par, gcnt, &d_sc->best_seam, &d_sc->seam_roots));
  CU_TRY(cudaGetLastError());
  B2M_TRY(b2m_fetch_scalars(ctx));
  sr->K = ctx->h_scalars->seam_roots;
  sr->par = par; sr->gflag = gflag;
  const unsigned long long best = ctx->h_scalars->best_seam;
  if (best) {
    const unsigned id = 0x3fffffffu - (unsigned)(best & 0x3fffffffull);
    seam_ent e;
    CU_TRY(cudaMemcpyAsync(&e, all_ent + id, sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    int r = 0;
    while (r + 1 < W && offs[r + 1] <= id) r++;
    sr->best_size = best >> 30;
    sr->best_rank = r;
    sr->best_slot = e.slot;
    sr->gstar = (long long)id;
  }
  return B2M_OK;
}

static int cc_seams(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const uint32_t *bits_own, const uint32_t *below,
                    const cc_geom &cg, cc_nodes nodes, int conn, b2m_scalars *d_sc, seam_result *sr) {
  static const bool force_slow = getenv("B2M_SEAM_SLOW") && atoi(getenv("B2M_SEAM_SLOW")) > 0;
  if (!force_slow) {
    int over = 0;
    B2M_TRY(cc_seams_fast(ctx, comm, sl, bits_own, below, cg, nodes, conn, d_sc, sr, &over));
    if (!over) return B2M_OK;
  }
  // the slow path needs the scalars of the local labelling on the host as well (nroots / best): its first
  // b2m_sync_scalars brings them
  return cc_seams_slow(ctx, comm, sl, bits_own, below, cg, nodes, conn, d_sc, sr);
}

// CC part of the front: fills fo->fill / fo->keep and the raw bright bbox in d_sc->lo/hi.
// g = geometry of the EXT planes of this rank; bit rows are EXT buffers.
int b2m_cc_run(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const b2m_geom &g, const b2m_opts *o, b2m_scalars *d_sc,
               b2m_front_out *fo) {
  const size_t pw = (size_t)g.ny * g.w;                       // words per plane
  const long long own_words = (long long)pw * sl.nzl;
  cc_geom cg = {g.nx, g.ny, sl.nzl, g.w, own_words, sl.z0 == 0, sl.z0 + sl.nzl == sl.gnz};
  size_t wbytes = (size_t)g.nwords * 4;
  unsigned blocks = b2m_cdiv(own_words, 256);
  const bool cc = o->only_largest || o->fill_bubbles;
  const bool slabs = sl.world > 1;
  // bwlabelCore() refuses volumes narrower than 2 voxels in x or y (src/bwlabel.c:434-437: it prints a message and
  // leaves the mask as thresholded), so -b finds no bubbles and -l keeps every bright voxel
  const bool refused = g.nx < 2 || g.ny < 2;
  B2M_TRY(b2m_reserve(ctx, BUF_FG, wbytes));
  uint32_t *fg = b2m_ptr<uint32_t>(ctx, BUF_FG);
  uint32_t *bg = nullptr;
  if (o->fill_bubbles) {
    B2M_TRY(b2m_reserve(ctx, BUF_BG, wbytes));
    bg = b2m_ptr<uint32_t>(ctx, BUF_BG);
  }
  B2M_TRY(b2m_reserve(ctx, BUF_MB, wbytes));
  uint32_t *mb = b2m_ptr<uint32_t>(ctx, BUF_MB);
  const int classic = o->backend == B2M_BACKEND_CLASSIC;
  // all EXT planes: the halo bits equal the neighbour's.  Skipped when the smooth wrote the same rows for this isolevel
  if (!(fo->bits_ready && !fo->iso_reset && fo->bits_iso == fo->iso))
    B2M_TRY(b2m_threshold_run(ctx, fo->S, g, fo->iso, fg, bg, mb, classic));
  fo->fill = nullptr;
  fo->keep = nullptr;
  const uint32_t *bright = fg;
  cc_nodes nodes = {nullptr, (uint32_t)own_words};
  const size_t own_off = (size_t)sl.hl * pw;                   // first own word in an EXT bit buffer
  if (cc && !refused) {
    if ((unsigned long long)own_words > (1ull << 27)) {  // run slots = word * 16 + run stay below 2^31 (bit 31 tags seam roots)
      b2m_set_error("slab of more than 2^27 bit words (2^32 voxels): %lld words", own_words);
      return B2M_EARG;
    }
    B2M_TRY(b2m_reserve(ctx, BUF_NODES, (size_t)own_words * 16 * sizeof(uint2)));
    nodes.p = b2m_ptr<uint2>(ctx, BUF_NODES);
  }
  // one halo plane of an EXT bit buffer from each neighbour (own boundary planes go the other way)
  auto halo_bits = [&](uint32_t *ext) -> int {
    if (!slabs) return B2M_OK;
    return b2m_comm_exchange(ctx, comm, ext + own_off + (size_t)(sl.nzl - 1) * pw, sl.hh ? pw * 4 : 0, ext, sl.hl ? pw * 4 : 0,
                             ext + own_off, sl.hl ? pw * 4 : 0, ext + own_off + (size_t)sl.nzl * pw, sl.hh ? pw * 4 : 0);
  };
  if (o->fill_bubbles && refused) { fo->fill = fg; }  // no bubbles found: the "filled" mask is the thresholded one
  if (o->fill_bubbles && !refused) {
    B2M_TRY(b2m_reserve(ctx, BUF_FILL, wbytes));
    uint32_t *fill = b2m_ptr<uint32_t>(ctx, BUF_FILL);
    B2M_TRY(cc_label(ctx, bg + own_off, cg, nodes, 6, nullptr, &d_sc->nroots_bg, d_sc));
    long long nroots_ovr = -1;
    if (slabs) {
      seam_result sr;
      B2M_TRY(cc_seams(ctx, comm, sl, bg + own_off, bg, cg, nodes, 6, d_sc, &sr));
      unsigned long long tot = 0;
      for (int r = 0; r < sl.world; r++) tot += b2m_sc(ctx, comm, r)->nroots_bg;
      nroots_ovr = (long long)(tot - (sr.M - sr.K));
      B2M_TRY(cc_seams_apply(ctx, sl, nodes, &sr, 1, false));
    }
    KT_LAUNCH(ctx, "cc_select", k_cc_select<<<blocks, 256, 0, ctx->stream>>>(bg + own_off, own_words, nodes, 1, nullptr, &d_sc->nroots_bg, fg + own_off, fill + own_off, -2, nroots_ovr));
    B2M_TRY(halo_bits(fill));
    fo->fill = fill;
    bright = fill;
  }
  uint32_t *largest = nullptr, *keep = nullptr;
  if (o->only_largest) {
    B2M_TRY(b2m_reserve(ctx, BUF_LARGEST, wbytes));
    B2M_TRY(b2m_reserve(ctx, BUF_KEEP, wbytes));
    largest = b2m_ptr<uint32_t>(ctx, BUF_LARGEST);
    keep = b2m_ptr<uint32_t>(ctx, BUF_KEEP);
    if (refused) CU_TRY(cudaMemcpyAsync(largest, bright, wbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    else {
    B2M_TRY(cc_label(ctx, bright + own_off, cg, nodes, 18, &d_sc->best_fg, &d_sc->nroots_fg, d_sc));
    long long sel = -2;
    if (slabs) {
      seam_result sr;
      B2M_TRY(cc_seams(ctx, comm, sl, bright + own_off, bright, cg, nodes, 18, d_sc, &sr));
      // global winner: most voxels, ties to the earliest first voxel = lowest (rank, slot) (src/bwlabel.c:462-466)
      unsigned long long bsize = 0; int brank = -1; unsigned bslot = 0;
      for (int r = 0; r < sl.world; r++) {
        const unsigned long long k = b2m_sc(ctx, comm, r)->best_fg;
        if (!k) continue;
        const unsigned long long sz = k >> 31;
        const unsigned slot = 0x7fffffffu - (unsigned)(k & 0x7fffffffull);
        if (sz > bsize) { bsize = sz; brank = r; bslot = slot; }
      }
      bool win = false;
      if (sr.best_size) {
        const unsigned long long sz = sr.best_size;
        if (sz > bsize || (sz == bsize && (sr.best_rank < brank || (sr.best_rank == brank && sr.best_slot <= bslot)))) {
          win = true;
          brank = -1;
        }
      }
      sel = (brank == sl.rank) ? (long long)bslot : -1;
      B2M_TRY(cc_seams_apply(ctx, sl, nodes, &sr, 0, win));
    }
    KT_LAUNCH(ctx, "cc_select", k_cc_select<<<blocks, 256, 0, ctx->stream>>>(bright + own_off, own_words, nodes, 0, &d_sc->best_fg, nullptr, nullptr, largest + own_off, sel, -1));
    B2M_TRY(halo_bits(largest));
    }
    fo->keep = keep;
  }
  {
    // global-z indexing: EXT buffers shifted down by ez0 planes
    const long long shift = (long long)sl.ez0 * (long long)pw;
    cc_geom dg = {g.nx, g.ny, sl.gnz, g.w, 0, 1, 1};
    dim3 dgrid(b2m_cdiv((size_t)g.ny * g.w, 256), b2m_cdiv(sl.nzl, DIL_ZC));
    ibits_params ip;
    ip.mb = mb - shift;
    ip.classic = classic;
    if (classic) {
      ip.cM = fo->vmin < fo->iso ? 0xffffffffu : 0u;
      ip.cE = fo->edge_max < fo->iso ? 0xffffffffu : 0u;
    } else {
      const float dm = fo->vmin - fo->iso, de = fo->edge_max - fo->iso;  // f32 subtractions, as mc_inside_at
      ip.cM = dm > -FLT_EPSILON ? 0xffffffffu : 0u;
      ip.cE = de > -FLT_EPSILON ? 0xffffffffu : 0u;
    }
    KT_LAUNCH(ctx, "dilate_bbox", k_dilate_bbox<<<dgrid, 256, 0, ctx->stream>>>(largest ? largest - shift : nullptr, bright - shift, dg, sl.z0, sl.z0 + sl.nzl,
                                                                              keep ? keep - shift : nullptr, d_sc->lo, ip));
    if (keep) B2M_TRY(halo_bits(keep));
    B2M_TRY(halo_bits(mb));
    fo->ibits = mb;
  }
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}
