// weld.cu — kernel 4: vertex weld (+ degenerate-triangle removal) on the GPU.
//
// Reference behaviour (/root/reference/src/meshify.c:45-106 unify_vertices, :113-168
// remove_degenerate_triangles, src/radixsort.c:277-348):
//   key_i = (float) |p_i - p_0|  (FP64 distance rounded to f32); stable sort by key; greedy scan in
//   key order: an unassigned vertex becomes a head; every later vertex whose f32 key difference to
//   the head is < 1e-5f and whose FP64 distance to the head is < 1e-5f is (re)assigned to it; the
//   merged vertex takes the coordinates of the highest original index of its cluster.  Then every
//   triangle with a repeated vertex or a Kahan/Heron area below FLT_EPSILON is dropped.
//
// Here vertices are born unique per grid edge (mc.cu), so the reference procedure only has to run on
// the ITEMS flagged by the emit pass: vertices within 2e-5 of a grid corner — the only way two
// vertices of different grid edges can be closer than the tolerance.
//   Lewiner: one item per flagged vertex, original index = vertex id.
//   Classic: the reference welds a triangle SOUP (src/oldcubes.c:505-519), so ties between equal f32
//     keys are broken by SOUP index and one grid-edge vertex can be split between two clusters
//     (copies sorted before / after a later head).  Items are therefore the flagged soup COPIES
//     (original index = 3*triangle + corner, position = the FP64 variant of the emitting cube).
// The item set is tiny (~1e-3 of the vertices).  It is sorted by original index, keyed, sorted by
// (key, original index), split wherever consecutive keys are >= tol apart (the reference's window
// never crosses such a gap) and the reference's greedy loop runs per segment.  Every cluster becomes
// one output vertex: the slot of the vertex id of its top (highest original index) item, or an
// extra slot appended after the last vertex when that id already owns a cluster (classic splits).
// Vertex ids that own no cluster are dead.  Index compaction and item lookup use two block tables
// (one u32 per 1024 ids) instead of per-vertex remap arrays: the triangle pass reads ~0.3 MB of
// cache-resident tables instead of gathering from two nv-sized arrays.
//
// Slabs (multi-GPU): the items of all ranks are all-gathered and every rank resolves the same item
// set; vertex ids are global, each rank compacts its own vertex block and its own triangles.
#include "common.cuh"

#define WB_SHIFT 10 /* ids per block-table entry */

__device__ __forceinline__ double dist_rn(const double *a, const double *b) {  // dx(): sqrt(sqr+sqr+sqr), no FMA
  double x = __dsub_rn(a[0], b[0]), y = __dsub_rn(a[1], b[1]), z = __dsub_rn(a[2], b[2]);
  double s = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
  return __dsqrt_rn(s);
}

// ---- item preparation --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_w_items_all(const double *__restrict__ verts, unsigned n, b2m_item *__restrict__ items) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  b2m_item it;
  it.pos[0] = verts[3 * (size_t)i]; it.pos[1] = verts[3 * (size_t)i + 1]; it.pos[2] = verts[3 * (size_t)i + 2];
  it.id = i; it.vid = i;
  items[i] = it;
}

// All sort keys are (major << ib) | minor with ib = bits needed for an item index, so that the radix
// sort runs over 32+ib bits instead of 64.
__global__ void __launch_bounds__(256) k_w_key1(const b2m_item *__restrict__ items, unsigned n, int ib, uint64_t *__restrict__ keys) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((uint64_t)items[i].id << ib) | (uint64_t)i;
}

// items in original-index order: S_id / S_vid / S_pos, and the (key, rank) sort keys
__global__ void __launch_bounds__(256) k_w_build(const b2m_item *__restrict__ items, const uint64_t *keys1, unsigned n, int ib,
                                                 const double *__restrict__ p0, uint32_t *__restrict__ S_id,
                                                 uint32_t *__restrict__ S_vid, double *__restrict__ S_pos, uint64_t *keys2) {
  unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const b2m_item it = items[(uint32_t)(keys1[r] & ((1ull << ib) - 1ull))];
  S_id[r] = it.id;
  S_vid[r] = it.vid;
  S_pos[3 * (size_t)r] = it.pos[0]; S_pos[3 * (size_t)r + 1] = it.pos[1]; S_pos[3 * (size_t)r + 2] = it.pos[2];
  double o[3] = {p0[0], p0[1], p0[2]};
  float k = (float)dist_rn(o, it.pos);
  keys2[r] = ((uint64_t)__float_as_uint(k) << ib) | (uint64_t)r;  // keys are >= 0: bit order == float order
}

// one thread per sorted position; the thread at the start of a segment runs the reference's greedy
// loop over its segment.  head[j] = sorted position of the head item j ended up with.
__global__ void __launch_bounds__(128) k_weld_resolve(const double *__restrict__ S_pos, const uint64_t *__restrict__ keys,
                                                      unsigned n, int ib, uint32_t *__restrict__ head) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tol = 0.00001f;
  const double told = (double)tol;
  const uint64_t lowm = (1ull << ib) - 1ull;
#define KEYF(q_) __uint_as_float((uint32_t)(keys[q_] >> ib))
#define KEYR(q_) ((uint32_t)(keys[q_] & lowm))
  float ki = KEYF(i);
  if (i > 0) {
    if (__fsub_rn(ki, KEYF(i - 1)) < tol) return;  // not a segment start
  }
  unsigned end = i + 1;
  while (end < n) {
    if (!(__fsub_rn(KEYF(end), KEYF(end - 1)) < tol)) break;
    end++;
  }
  for (unsigned q = i; q < end; q++) head[q] = 0xffffffffu;
  for (unsigned q = i; q < end; q++) {
    if (head[q] != 0xffffffffu) continue;  // already assigned
    const float k0 = KEYF(q);
    const double *pq = S_pos + 3 * (size_t)KEYR(q);
    for (unsigned j = q; j < end; j++) {
      if (!(__fsub_rn(KEYF(j), k0) < tol)) break;
      if (dist_rn(pq, S_pos + 3 * (size_t)KEYR(j)) < told) head[j] = q;
    }
  }
#undef KEYF
#undef KEYR
}

// top[h] = highest original-index rank in cluster h (src/meshify.c:99-100); inv[rank] = sorted position
__global__ void __launch_bounds__(256) k_weld_top(const uint64_t *__restrict__ keys, unsigned n, int ib, const uint32_t *__restrict__ head,
                                                  uint32_t *__restrict__ top, uint32_t *__restrict__ inv) {
  unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t r = (uint32_t)(keys[j] & ((1ull << ib) - 1ull));
  atomicMax(&top[head[j]], r);
  inv[r] = j;
}

// cluster ownership keys: heads -> (owner vertex id << 32 | head position), everything else -> ~0
__global__ void __launch_bounds__(256) k_w_key3(const uint32_t *__restrict__ head, const uint32_t *__restrict__ top,
                                                const uint32_t *__restrict__ S_vid, unsigned n, int ib, uint64_t *__restrict__ keys3) {
  unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  keys3[j] = head[j] == j ? (((uint64_t)S_vid[top[j]] << ib) | (uint64_t)j) : (1ull << (32 + ib)) - 1ull;  // non-heads sort last
}
// sorted clusters: the first cluster of an owner id is primary (takes the id's slot), later ones are extras
__global__ void __launch_bounds__(256) k_w_cluster_flags(const uint64_t *__restrict__ keys3, unsigned n, int ib, uint32_t *__restrict__ xflag,
                                                         uint32_t *__restrict__ pflag) {
  unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint64_t k = keys3[q];
  const bool valid = k != (1ull << (32 + ib)) - 1ull;
  const bool primary = valid && (q == 0 || (uint32_t)(keys3[q - 1] >> ib) != (uint32_t)(k >> ib));
  xflag[q] = valid && !primary;
  pflag[q] = primary;
}
// cslot[head position] = owner id (primary) or 0x80000000 | extra rank; owners[] = sorted unique owner ids
__global__ void __launch_bounds__(256) k_w_cluster_slots(const uint64_t *__restrict__ keys3, unsigned n, int ib,
                                                         const uint32_t *__restrict__ xscan, const uint32_t *__restrict__ pscan,
                                                         uint32_t *__restrict__ cslot, uint32_t *__restrict__ owners) {
  unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint64_t k = keys3[q];
  if (k == (1ull << (32 + ib)) - 1ull) return;
  const uint32_t ov = (uint32_t)(k >> ib), h = (uint32_t)(k & ((1ull << ib) - 1ull));
  const bool primary = q == 0 || (uint32_t)(keys3[q - 1] >> ib) != ov;
  if (primary) { cslot[h] = ov; owners[pscan[q]] = ov; }
  else cslot[h] = 0x80000000u | xscan[q];
}

__device__ __forceinline__ unsigned lower_bound_u32(const uint32_t *__restrict__ a, unsigned n, uint32_t v) {
  unsigned lo = 0, hi = n;
  while (lo < hi) {
    unsigned mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_w_key4(const uint32_t *__restrict__ S_vid, unsigned n, uint64_t *__restrict__ keys4) {
  unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) keys4[r] = (uint64_t)S_vid[r];   // only the distinct ids matter
}
// sorted item vertex ids: a distinct id that owns no cluster is dead
__global__ void __launch_bounds__(256) k_w_dead_flags(const uint64_t *__restrict__ keys4, unsigned n, const uint32_t *__restrict__ owners,
                                                      const uint32_t *__restrict__ n_owners, uint32_t *__restrict__ dflag) {
  unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t v = (uint32_t)keys4[q];
  bool dead = false;
  if (q == 0 || (uint32_t)keys4[q - 1] != v) {
    const unsigned no = *n_owners;
    const unsigned p = lower_bound_u32(owners, no, v);
    dead = !(p < no && owners[p] == v);
  }
  dflag[q] = dead;
}
__global__ void __launch_bounds__(256) k_w_dead_list(const uint64_t *__restrict__ keys4, unsigned n, const uint32_t *__restrict__ dscan,
                                                     const uint32_t *__restrict__ n_dead, uint32_t *__restrict__ R_vid) {
  unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t me = dscan[q], next = q + 1 < n ? dscan[q + 1] : *n_dead;
  if (next != me) R_vid[me] = (uint32_t)keys4[q];
}

// Lewiner: every item is its own vertex id (ids are unique and already in id order), so every cluster is
// primary, owned by the id of its top item, and an item is dead iff it is not the top of its cluster.
__global__ void __launch_bounds__(256) k_w_unique_slots(unsigned n, const uint32_t *__restrict__ inv, const uint32_t *__restrict__ head,
                                                        const uint32_t *__restrict__ top, const uint32_t *__restrict__ S_vid,
                                                        uint32_t *__restrict__ cslot, uint32_t *__restrict__ dflag) {
  unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t j = inv[r], h = head[j];
  if (h == j) cslot[j] = S_vid[top[j]];
  dflag[r] = top[h] != r;
}
__global__ void __launch_bounds__(256) k_w_unique_dead_list(unsigned n, const uint32_t *__restrict__ S_vid, const uint32_t *__restrict__ dscan,
                                                            const uint32_t *__restrict__ n_dead, uint32_t *__restrict__ R_vid) {
  unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t me = dscan[r], next = r + 1 < n ? dscan[r + 1] : *n_dead;
  if (next != me) R_vid[me] = S_vid[r];
}

// block tables: T[b] = number of list entries < b << WB_SHIFT
__global__ void __launch_bounds__(256) k_w_table(const uint32_t *__restrict__ list, const uint32_t *__restrict__ n_ptr, unsigned n_fixed,
                                                 uint32_t *__restrict__ T, unsigned nb) {
  unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const unsigned n = n_ptr ? *n_ptr : n_fixed;
  const uint64_t lim = (uint64_t)b << WB_SHIFT;
  T[b] = lim > 0xffffffffull ? n : lower_bound_u32(list, n, (uint32_t)lim);
}

struct weld_tables {
  const uint32_t *Q, *S_id;      // item lookup by original index (null when there are no items)
  const uint32_t *P, *R_vid;     // dead vertex ids (null when nothing died)
  const uint32_t *S_out, *S_top; // per item: output vertex index, rank of the cluster's top item
  const double *S_pos;
};
// number of dead ids below g; *dead = g itself is dead
__device__ __forceinline__ uint32_t dead_below(const weld_tables &w, uint32_t g, bool *dead) {
  *dead = false;
  if (!w.P) return 0;
  const uint32_t b = g >> WB_SHIFT, lo = __ldg(w.P + b), hi = __ldg(w.P + b + 1);
  uint32_t c = lo;
  for (uint32_t k = lo; k < hi; k++) {
    const uint32_t v = __ldg(w.R_vid + k);
    if (v < g) c++;
    else { *dead = v == g; break; }
  }
  return c;
}
__device__ __forceinline__ int find_item(const weld_tables &w, uint32_t id) {
  if (!w.Q) return -1;
  const uint32_t b = id >> WB_SHIFT, lo = __ldg(w.Q + b), hi = __ldg(w.Q + b + 1);
  for (uint32_t k = lo; k < hi; k++) {
    const uint32_t v = __ldg(w.S_id + k);
    if (v == id) return (int)k;
    if (v > id) break;
  }
  return -1;
}

// per item: output vertex index of its cluster and the cluster's top item
__global__ void __launch_bounds__(256) k_w_item_out(unsigned n, const uint32_t *__restrict__ inv, const uint32_t *__restrict__ head,
                                                    const uint32_t *__restrict__ top, const uint32_t *__restrict__ cslot,
                                                    weld_tables w, unsigned NV, const uint32_t *__restrict__ n_dead,
                                                    uint32_t *__restrict__ S_out, uint32_t *__restrict__ S_top) {
  unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t h = head[inv[r]];
  const uint32_t slot = cslot[h];
  uint32_t out;
  if (slot & 0x80000000u) out = NV - *n_dead + (slot & 0x7fffffffu);
  else { bool d; out = slot - dead_below(w, slot, &d); }
  S_out[r] = out;
  S_top[r] = top[h];
}

// geometry of one rank's vertex blocks in the global (pre-weld) numbering
struct weld_geom {
  unsigned e_off, nv_edge;   // own edge vertices: global ids e_off .. e_off+nv_edge
  unsigned c_base, nv_c;     // own centroid vertices: global ids c_base .. c_base+nv_c  (c_base = NVE + c_off)
  unsigned halo0, halo1;     // next rank's first-plane vertices (positions in halo[])
  unsigned t_off;            // global index of own triangle 0
  unsigned NV;               // global vertex count before the weld
  int last_rank, classic_soup;
};
struct weld_bases {  // computed on the device by k_w_bases, read back once
  uint32_t new_e_off, nve_new, new_c_base, nvc_new, n_dead, n_extra, pad[2];
};
__global__ void k_w_bases(weld_geom g, weld_tables w, const uint32_t *n_dead, const uint32_t *n_extra, weld_bases *out, b2m_scalars *sc) {
  bool d;
  const uint32_t a = g.e_off - dead_below(w, g.e_off, &d);
  const uint32_t b = (g.e_off + g.nv_edge) - dead_below(w, g.e_off + g.nv_edge, &d);
  const uint32_t c = g.c_base - dead_below(w, g.c_base, &d);
  const uint32_t e = (g.c_base + g.nv_c) - dead_below(w, g.c_base + g.nv_c, &d);
  out->new_e_off = a; out->nve_new = b - a; out->new_c_base = c; out->nvc_new = e - c;
  out->n_dead = n_dead ? *n_dead : 0; out->n_extra = n_extra ? *n_extra : 0;
  out->pad[0] = out->pad[1] = 0;
  // the host reads them with the last sync of the call (b2m_weld_finish)
  sc->wb_new_e_off = a; sc->wb_nve_new = b - a; sc->wb_new_c_base = c; sc->wb_nvc_new = e - c;
  sc->wb_n_dead = out->n_dead; sc->wb_n_extra = out->n_extra;
}

// own vertices -> compacted local array [edge block | centroid block | extras (last rank)]
__global__ void __launch_bounds__(256) k_compact_verts(const double *__restrict__ vin, double *__restrict__ vout, weld_geom g,
                                                       weld_tables w, const weld_bases *__restrict__ bs) {
  unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.nv_edge + g.nv_c) return;
  const bool edge = j < g.nv_edge;
  const uint32_t gid = edge ? g.e_off + j : g.c_base + (j - g.nv_edge);
  bool dead;
  const uint32_t nid = gid - dead_below(w, gid, &dead);
  if (dead) return;
  const size_t o = 3 * (size_t)(edge ? nid - bs->new_e_off : bs->nve_new + (nid - bs->new_c_base));
  vout[o] = vin[3 * (size_t)j]; vout[o + 1] = vin[3 * (size_t)j + 1]; vout[o + 2] = vin[3 * (size_t)j + 2];
}
// cluster heads write the coordinates of their top item into the cluster's output slot (classic: the
// surviving soup copy's FP64 variant; extras: the split-off vertices)
__global__ void __launch_bounds__(256) k_w_patch(unsigned n, const uint32_t *__restrict__ head, const uint32_t *__restrict__ top,
                                                 const uint32_t *__restrict__ cslot, weld_geom g, weld_tables w,
                                                 const weld_bases *__restrict__ bs, double *__restrict__ vout) {
  unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || head[j] != j) return;
  const uint32_t slot = cslot[j];
  size_t o;
  if (slot & 0x80000000u) {
    if (!g.last_rank) return;
    o = (size_t)bs->nve_new + bs->nvc_new + (slot & 0x7fffffffu);
  } else {
    bool d;
    if (slot - g.e_off < g.nv_edge) o = (size_t)(slot - dead_below(w, slot, &d)) - bs->new_e_off;
    else if (slot - g.c_base < g.nv_c) o = (size_t)bs->nve_new + ((slot - dead_below(w, slot, &d)) - bs->new_c_base);
    else return;  // another rank's vertex
  }
  const double *s = w.S_pos + 3 * (size_t)top[j];
  vout[3 * o] = s[0]; vout[3 * o + 1] = s[1]; vout[3 * o + 2] = s[2];
}

// Triangle clean-up (src/meshify.c:113-168) in two streaming passes over the index array:
//   k_tri_degen : keep bit per triangle (+ count per 32) - the reference's FP64 needle test, run only for the triangles
//                 that can fail it (a vertex within B2M_NEAR_TOL of a grid corner, a centroid vertex, a vertex of the
//                 next rank: see B2M_NEAR_TOL); the others never touch a vertex position;
//   k_tri_finish: kept triangles -> welded vertex numbering, written at their compacted position.
// welded index and (optionally) position of one triangle corner; slow = the vertex may own a weld item
__device__ __forceinline__ int tri_corner(const weld_geom &g, const weld_tables &w, uint32_t vid, uint32_t soup_id, bool slow,
                                          const double *__restrict__ verts, const double *__restrict__ halo, const double **src,
                                          unsigned int *__restrict__ overflow) {
  const int k = slow ? find_item(w, g.classic_soup ? soup_id : vid) : -1;
  if (k >= 0) {
    if (src) *src = w.S_pos + 3 * (size_t)__ldg(w.S_top + k);
    return (int)__ldg(w.S_out + k);
  }
  bool dead;
  const int idx = (int)(vid - dead_below(w, vid, &dead));
  if (src) {
    if (vid - g.e_off < g.nv_edge) *src = verts + 3 * (size_t)(vid - g.e_off);
    else if (vid - g.c_base < g.nv_c) *src = verts + 3 * (size_t)(g.nv_edge + (vid - g.c_base));
    else if (vid - g.halo0 < g.halo1 - g.halo0) *src = halo + 3 * (size_t)(vid - g.halo0);
    else { atomicOr(overflow, 2u); *src = verts; }
  }
  return idx;
}
// may this vertex fail the needle test / own a weld item?  (own edge vertices: the flag bit; everything else: yes)
__device__ __forceinline__ bool vert_slow(const weld_geom &g, const uint32_t *__restrict__ nearbits, uint32_t vid) {
  const uint32_t l = vid - g.e_off;
  if (!nearbits || l >= g.nv_edge) return true;
  return (__ldg(nearbits + (l >> 5)) >> (l & 31u)) & 1u;
}

// Both passes: a block of 256 threads covers 1024 consecutive triangles, thread t taking t, t+256, t+512, t+768 (twelve
// index loads in flight per thread; a warp's ballot is the keep / slow word of 32 consecutive triangles).
#define TRI_PER_THREAD 4
#define TRI_PER_BLOCK (256 * TRI_PER_THREAD)
/* measured on G1024: 0.402 ms at 8 CTAs per SM (32 registers), 0.445 at 6, 0.483 unconstrained (48 registers) */
#ifndef TD_MINB
#define TD_MINB 8
#endif
__global__ void __launch_bounds__(256, TD_MINB) k_tri_degen(const int *__restrict__ tris, unsigned nt, const double *__restrict__ verts,
                                                   const double *__restrict__ halo, const uint32_t *__restrict__ nearbits,
                                                   weld_geom g, weld_tables w, uint32_t *__restrict__ keepbits,
                                                   uint32_t *__restrict__ keepcnt, uint32_t *__restrict__ slowbits,
                                                   unsigned int *__restrict__ overflow) {
  // the triangles that need the test are gathered per block so that whole warps run it (a few percent of the
  // triangles, spread evenly: without the gather nearly every warp would walk the slow path for one or two lanes)
  __shared__ unsigned s_n;
  __shared__ unsigned short s_list[TRI_PER_BLOCK];
  __shared__ unsigned char s_keep[TRI_PER_BLOCK];
  const unsigned tid = threadIdx.x, lane = tid & 31u, b0 = blockIdx.x * TRI_PER_BLOCK;
  if (tid == 0) s_n = 0;
  __syncthreads();
  uint32_t v[TRI_PER_THREAD][3];
#pragma unroll
  for (int k = 0; k < TRI_PER_THREAD; k++) {
    const unsigned i = b0 + k * 256 + tid;
    const int *t = tris + 3 * (size_t)(i < nt ? i : 0);
    v[k][0] = (uint32_t)__ldg(t); v[k][1] = (uint32_t)__ldg(t + 1); v[k][2] = (uint32_t)__ldg(t + 2);
  }
  unsigned slowm = 0;
#pragma unroll
  for (int k = 0; k < TRI_PER_THREAD; k++) {
    const unsigned i = b0 + k * 256 + tid;
    const bool slow = i < nt && (vert_slow(g, nearbits, v[k][0]) | vert_slow(g, nearbits, v[k][1]) | vert_slow(g, nearbits, v[k][2]));
    const unsigned m = __ballot_sync(0xffffffffu, slow);
    unsigned base = 0;
    if (lane == 0 && m) base = atomicAdd(&s_n, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (slow) s_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(k * 256 + tid);
    if (lane == 0 && i < nt) slowbits[i >> 5] = m;
    slowm |= (slow ? 1u : 0u) << k;
  }
  __syncthreads();
  const unsigned ns = s_n;
  for (unsigned q = tid; q < ns; q += blockDim.x) {
    const unsigned lt = s_list[q], ti = b0 + lt;
    double p[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const uint32_t vid = (uint32_t)tris[3 * (size_t)ti + c];
      const double *src;
      tri_corner(g, w, vid, 3u * (g.t_off + ti) + (uint32_t)c, true, verts, halo, &src, overflow);
      p[c][0] = src[0]; p[c][1] = src[1]; p[c][2] = src[2];
    }
    double l = dist_rn(p[0], p[1]), m = dist_rn(p[0], p[2]), n = dist_rn(p[1], p[2]);
    double cc = fmin(fmin(l, m), n), aa = fmax(fmax(l, m), n);
    double bb = __dsub_rn(__dsub_rn(__dadd_rn(__dadd_rn(l, m), n), aa), cc);
    double amb = __dsub_rn(aa, bb);
    double t1 = __dsub_rn(cc, amb);
    unsigned char keep = 1;
    if (t1 <= 0.0) keep = 0;
    else {
      double prod = __dmul_rn(__dmul_rn(__dmul_rn(__dadd_rn(aa, __dadd_rn(bb, cc)), t1), __dadd_rn(cc, amb)),
                              __dadd_rn(aa, __dsub_rn(bb, cc)));
      double area4 = __dmul_rn(0.25, __dsqrt_rn(prod));
      if (area4 < (double)FLT_EPSILON) keep = 0;
    }
    s_keep[lt] = keep;
  }
  __syncthreads();
  // keep flags leave the kernel as one bit per triangle + a count per 32 triangles (the compaction offsets come
  // from a scan over nt/32 counts instead of nt flags)
#pragma unroll
  for (int k = 0; k < TRI_PER_THREAD; k++) {
    const unsigned i = b0 + k * 256 + tid;
    const bool keep = i < nt && (!((slowm >> k) & 1u) || s_keep[k * 256 + tid]);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0 && i < nt) { keepbits[i >> 5] = m; keepcnt[i >> 5] = (uint32_t)__popc(m); }
  }
}

__global__ void __launch_bounds__(256) k_tri_finish(const int *__restrict__ tin, int *__restrict__ tout,
                                                    const uint32_t *__restrict__ keepbits, const uint32_t *__restrict__ cnt_scanned,
                                                    const uint32_t *__restrict__ slowbits, unsigned nt, weld_geom g, weld_tables w) {
  const unsigned tid = threadIdx.x, bit = tid & 31u, b0 = blockIdx.x * TRI_PER_BLOCK;
  uint32_t v[TRI_PER_THREAD][3], km[TRI_PER_THREAD], sm[TRI_PER_THREAD], base[TRI_PER_THREAD];
#pragma unroll
  for (int k = 0; k < TRI_PER_THREAD; k++) {
    const unsigned i = b0 + k * 256 + tid;
    const bool in = i < nt;
    const int *t = tin + 3 * (size_t)(in ? i : 0);
    v[k][0] = (uint32_t)__ldg(t); v[k][1] = (uint32_t)__ldg(t + 1); v[k][2] = (uint32_t)__ldg(t + 2);
    km[k] = in ? __ldg(keepbits + (i >> 5)) : 0u;
    sm[k] = in ? __ldg(slowbits + (i >> 5)) : 0u;
    base[k] = in ? __ldg(cnt_scanned + (i >> 5)) : 0u;
  }
#pragma unroll
  for (int k = 0; k < TRI_PER_THREAD; k++) {
    if (!((km[k] >> bit) & 1u)) continue;
    const unsigned i = b0 + k * 256 + tid;
    const uint32_t me = base[k] + (uint32_t)__popc(km[k] & ((1u << bit) - 1u));
    const bool slow = (sm[k] >> bit) & 1u;  // only such a triangle can hold a vertex that owns a weld item
    int idx[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
      idx[c] = (w.Q || w.P) ? tri_corner(g, w, v[k][c], 3u * (g.t_off + i) + (uint32_t)c, slow, nullptr, nullptr, nullptr, nullptr)
                            : (int)v[k][c];
    const size_t o = 3 * (size_t)me;
    tout[o] = idx[0]; tout[o + 1] = idx[1]; tout[o + 2] = idx[2];
  }
}

// workspace carving: one arena (BUF_WELD) for all the small per-item arrays
struct carve {
  char *p;
  size_t off;
  template <typename T> T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *r = reinterpret_cast<T *>(p + off);
    off += n * sizeof(T);
    return r;
  }
};

// Enqueues the whole weld + triangle clean-up without waiting for the device: every count the host needs afterwards
// (weld bases, surviving triangles, consistency flags) is left in the scalar block, which the caller fetches with ONE
// synchronisation (b2m_sync_scalars) before b2m_weld_finish() turns it into the b2m_weld_out.  Buffers whose exact size
// is only known on the device are reserved at their upper bounds.
int b2m_weld_run(b2m_ctx *ctx, b2m_comm *comm, b2m_mesh_dev *mesh, int all_items, b2m_weld_out *wo) {
  b2m_scalars *d_sc = b2m_ptr<b2m_scalars>(ctx, BUF_SCALARS);
  const unsigned nvl = mesh->nv_edge + mesh->nv_c, nt = mesh->nt;
  cudaEvent_t e0 = ctx->ev[2 * B2M_T_WELD], e1 = ctx->ev[2 * B2M_T_WELD + 1];
  cudaEvent_t e2 = ctx->ev[2 * B2M_T_DEGEN], e3 = ctx->ev[2 * B2M_T_DEGEN + 1];
  CU_TRY(cudaEventRecord(e0, ctx->stream));
  weld_geom g;
  g.e_off = mesh->e_off; g.nv_edge = mesh->nv_edge;
  g.c_base = mesh->NVE + mesh->c_off; g.nv_c = mesh->nv_c;
  g.halo0 = mesh->halo0; g.halo1 = mesh->halo1;
  g.t_off = mesh->t_off;
  g.NV = mesh->NVE + mesh->NVC;
  g.last_rank = b2m_comm_rank(comm) == b2m_comm_world(comm) - 1;
  g.classic_soup = mesh->classic_soup;
  weld_tables w;
  memset(&w, 0, sizeof(w));
  memset(wo, 0, sizeof(*wo));

  // ---- the item set (all ranks' items, identical on every rank) ----
  b2m_item *items = nullptr;
  unsigned n = 0;
  if (all_items) {
    n = nvl;
    B2M_TRY(b2m_reserve(ctx, BUF_CAND, (size_t)n * sizeof(b2m_item)));
    items = b2m_ptr<b2m_item>(ctx, BUF_CAND);
    KT_LAUNCH(ctx, "weld_items", k_w_items_all<<<b2m_cdiv(n, 256), 256, 0, ctx->stream>>>(mesh->verts, n, items));
  } else {
    B2M_TRY(b2m_comm_gather_items(ctx, comm, mesh->nitems, &items, &n));
  }
  double *verts = mesh->verts;
  wo->verts = verts;
  wo->n_items = n;
  if (n >= 2) {
    const unsigned nbQ = (unsigned)((((uint64_t)(mesh->classic_soup ? 3ull * mesh->NT : g.NV)) >> WB_SHIFT) + 2);
    const unsigned nbP = (unsigned)(((uint64_t)g.NV >> WB_SHIFT) + 2);
    size_t need = (size_t)n * (8 + 4 + 4 + 24 + 4 * 10) + (size_t)(nbQ + nbP) * 4 + 64 * 256;
    B2M_TRY(b2m_reserve(ctx, BUF_WELD, need));
    carve cv = {b2m_ptr<char>(ctx, BUF_WELD), 0};
    uint64_t *keys = cv.take<uint64_t>(n);
    uint32_t *S_id = cv.take<uint32_t>(n), *S_vid = cv.take<uint32_t>(n);
    double *S_pos = cv.take<double>(3 * (size_t)n);
    uint32_t *head = cv.take<uint32_t>(n), *top = cv.take<uint32_t>(n), *inv = cv.take<uint32_t>(n);
    uint32_t *cslot = cv.take<uint32_t>(n), *owners = cv.take<uint32_t>(n), *R_vid = cv.take<uint32_t>(n);
    uint32_t *f0 = cv.take<uint32_t>(n + 4), *f1 = cv.take<uint32_t>(n + 4);
    uint32_t *S_out = cv.take<uint32_t>(n), *S_top = cv.take<uint32_t>(n);
    uint32_t *Q = cv.take<uint32_t>(nbQ), *P = cv.take<uint32_t>(nbP);
    weld_bases *d_bs = cv.take<weld_bases>(1);
    uint32_t *d_cnt = cv.take<uint32_t>(8);  // [0] n_extra [1] n_owners [2] n_dead
    const unsigned nb = b2m_cdiv(n, 256);
    int ib = 1;
    while ((1ull << ib) < (unsigned long long)n) ib++;
    CU_TRY(cudaMemsetAsync(d_cnt, 0, 32, ctx->stream));
    KT_LAUNCH(ctx, "weld_keys", k_w_key1<<<nb, 256, 0, ctx->stream>>>(items, n, ib, keys));
    B2M_TRY(b2m_sort_u64(ctx, keys, n, 32 + ib));
    KT_LAUNCH(ctx, "weld_keys", k_w_build<<<nb, 256, 0, ctx->stream>>>(items, keys, n, ib, mesh->d_p0, S_id, S_vid, S_pos, keys));
    B2M_TRY(b2m_sort_u64(ctx, keys, n, 32 + ib));
    KT_LAUNCH(ctx, "weld_resolve", k_weld_resolve<<<b2m_cdiv(n, 128), 128, 0, ctx->stream>>>(S_pos, keys, n, ib, head));
    CU_TRY(cudaMemsetAsync(top, 0, (size_t)n * 4, ctx->stream));
    KT_LAUNCH(ctx, "weld_top", k_weld_top<<<nb, 256, 0, ctx->stream>>>(keys, n, ib, head, top, inv));
    if (mesh->classic_soup) {
      // cluster ownership: primary / extra slots, sorted unique owner ids
      KT_LAUNCH(ctx, "weld_own", k_w_key3<<<nb, 256, 0, ctx->stream>>>(head, top, S_vid, n, ib, keys));
      B2M_TRY(b2m_sort_u64(ctx, keys, n, 32 + ib));
      KT_LAUNCH(ctx, "weld_own", k_w_cluster_flags<<<nb, 256, 0, ctx->stream>>>(keys, n, ib, f0, f1));
      B2M_TRY(b2m_exclusive_scan_u32(ctx, f0, f0, n, d_cnt + 0));
      B2M_TRY(b2m_exclusive_scan_u32(ctx, f1, f1, n, d_cnt + 1));
      KT_LAUNCH(ctx, "weld_own", k_w_cluster_slots<<<nb, 256, 0, ctx->stream>>>(keys, n, ib, f0, f1, cslot, owners));
      // dead vertex ids: distinct item ids that own no cluster
      KT_LAUNCH(ctx, "weld_dead", k_w_key4<<<nb, 256, 0, ctx->stream>>>(S_vid, n, keys));
      B2M_TRY(b2m_sort_u64(ctx, keys, n, 32));
      KT_LAUNCH(ctx, "weld_dead", k_w_dead_flags<<<nb, 256, 0, ctx->stream>>>(keys, n, owners, d_cnt + 1, f0));
      B2M_TRY(b2m_exclusive_scan_u32(ctx, f0, f0, n, d_cnt + 2));
      KT_LAUNCH(ctx, "weld_dead", k_w_dead_list<<<nb, 256, 0, ctx->stream>>>(keys, n, f0, d_cnt + 2, R_vid));
    } else {
      KT_LAUNCH(ctx, "weld_own", k_w_unique_slots<<<nb, 256, 0, ctx->stream>>>(n, inv, head, top, S_vid, cslot, f0));
      B2M_TRY(b2m_exclusive_scan_u32(ctx, f0, f0, n, d_cnt + 2));
      KT_LAUNCH(ctx, "weld_dead", k_w_unique_dead_list<<<nb, 256, 0, ctx->stream>>>(n, S_vid, f0, d_cnt + 2, R_vid));
    }
    KT_LAUNCH(ctx, "weld_table", k_w_table<<<b2m_cdiv(nbP, 256), 256, 0, ctx->stream>>>(R_vid, d_cnt + 2, 0, P, nbP));
    KT_LAUNCH(ctx, "weld_table", k_w_table<<<b2m_cdiv(nbQ, 256), 256, 0, ctx->stream>>>(S_id, nullptr, n, Q, nbQ));
    w.Q = Q; w.S_id = S_id; w.P = P; w.R_vid = R_vid; w.S_out = S_out; w.S_top = S_top; w.S_pos = S_pos;
    KT_LAUNCH(ctx, "weld_out", k_w_item_out<<<nb, 256, 0, ctx->stream>>>(n, inv, head, top, cslot, w, g.NV, d_cnt + 2, S_out, S_top));
    KT_LAUNCH(ctx, "weld_out", k_w_bases<<<1, 1, 0, ctx->stream>>>(g, w, d_cnt + 2, d_cnt + 0, d_bs, d_sc));
    CU_TRY(cudaGetLastError());
    // the compaction runs whether or not anything merged (with no dead ids and no extras it is a plain copy): that is
    // only known on the device here.  Upper bound of the output: every own vertex + every item as an extra
    B2M_TRY(b2m_reserve(ctx, BUF_VERTS2, ((size_t)nvl + (g.last_rank ? n : 0)) * 24));
    double *v2 = b2m_ptr<double>(ctx, BUF_VERTS2);
    if (nvl) KT_LAUNCH(ctx, "compact_verts", k_compact_verts<<<b2m_cdiv(nvl, 256), 256, 0, ctx->stream>>>(verts, v2, g, w, d_bs));
    if (mesh->classic_soup)
      KT_LAUNCH(ctx, "weld_patch", k_w_patch<<<nb, 256, 0, ctx->stream>>>(n, head, top, cslot, g, w, d_bs, v2));
    wo->verts = v2;
  }
  CU_TRY(cudaEventRecord(e1, ctx->stream));
  CU_TRY(cudaEventRecord(e2, ctx->stream));
  // ---- degenerate triangles ----
  wo->tris = mesh->tris;
  if (nt > 0) {
    const size_t nw32 = ((size_t)nt + 31) / 32;
    B2M_TRY(b2m_reserve(ctx, BUF_FLAGS, nw32 * 12 + 64));
    uint32_t *kb = b2m_ptr<uint32_t>(ctx, BUF_FLAGS), *kc = kb + nw32, *ks = kc + nw32;
    KT_LAUNCH(ctx, "tri_degen", k_tri_degen<<<b2m_cdiv(nt, TRI_PER_BLOCK), 256, 0, ctx->stream>>>(mesh->tris, nt, verts, mesh->halo_verts, mesh->nearbits, g, w, kb, kc, ks, &d_sc->overflow));
    B2M_TRY(b2m_exclusive_scan_u32(ctx, kc, kc, nw32, &d_sc->n_tri_kept));
    // drop + renumber in one pass; the output is reserved for every triangle (how many survive is on the device)
    B2M_TRY(b2m_reserve(ctx, BUF_TRIS2, (size_t)nt * 12));
    int *t2 = b2m_ptr<int>(ctx, BUF_TRIS2);
    KT_LAUNCH(ctx, "tri_finish", k_tri_finish<<<b2m_cdiv(nt, TRI_PER_BLOCK), 256, 0, ctx->stream>>>(mesh->tris, t2, kb, kc, ks, nt, g, w));
    wo->tris = t2;
  } else {
    CU_TRY(cudaMemsetAsync(&d_sc->n_tri_kept, 0, 4, ctx->stream));
  }
  CU_TRY(cudaEventRecord(e3, ctx->stream));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

// after the caller's b2m_sync_scalars(): the counts of this rank's welded blocks
int b2m_weld_finish(b2m_ctx *ctx, b2m_comm *comm, const b2m_mesh_dev *mesh, b2m_weld_out *wo) {
  const b2m_scalars *h = ctx->h_scalars;
  if (h->overflow & 2u) { b2m_set_error("weld: triangle references a vertex outside this rank's blocks"); return B2M_ECUDA; }
  const bool last_rank = b2m_comm_rank(comm) == b2m_comm_world(comm) - 1;
  const unsigned NV = mesh->NVE + mesh->NVC;
  if (wo->n_items >= 2) {
    wo->n_dead = h->wb_n_dead; wo->n_extra = h->wb_n_extra;
    wo->nve_local = h->wb_nve_new; wo->nvc_local = h->wb_nvc_new;
    wo->v_edge_off = h->wb_new_e_off; wo->v_c_off = h->wb_new_c_base;
  } else {
    wo->n_dead = wo->n_extra = 0;
    wo->nve_local = mesh->nv_edge; wo->nvc_local = mesh->nv_c;
    wo->v_edge_off = mesh->e_off; wo->v_c_off = mesh->NVE + mesh->c_off;
  }
  wo->nx_local = last_rank ? wo->n_extra : 0;
  wo->nv_local = wo->nve_local + wo->nvc_local + wo->nx_local;
  wo->nt_local = h->n_tri_kept;
  wo->nv_global = NV - wo->n_dead + wo->n_extra;
  return B2M_OK;
}
