// Plan construction: COO (int64, fp32) -> CSR + CSR^T (int32) + row-block union panels.
// One-time setup per Laplacian / remap matrix; done on the host and uploaded.
// Replaces the COO->CSR conversion torch.sparse.mm performs on every call
// (reference modules/layers.py:164,167,962).
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "dsw_internal.cuh"

namespace dsw {

std::atomic<int64_t> g_launches{0};
std::atomic<int> g_mix_mode{1};
std::atomic<int64_t> g_options[DSW_OPT_COUNT] = {};

static thread_local cudaError_t t_last_err = cudaSuccess;
void set_cuda_error(cudaError_t e) { t_last_err = e; }
cudaError_t last_cuda_error() { return t_last_err; }

struct HostCsr {
  int32_t n_rows = 0, n_cols = 0;
  std::vector<int32_t> rowptr, col;
  std::vector<float> val;
};

// Stable counting sort by row, ascending column inside a row, duplicates summed.
static HostCsr build_csr(int32_t n_rows, int32_t n_cols, int64_t nnz, const int64_t* r,
                         const int64_t* c, const float* v) {
  HostCsr out;
  out.n_rows = n_rows;
  out.n_cols = n_cols;
  std::vector<int64_t> cnt(static_cast<size_t>(n_rows) + 1, 0);
  for (int64_t i = 0; i < nnz; ++i) cnt[r[i] + 1]++;
  for (int32_t i = 0; i < n_rows; ++i) cnt[i + 1] += cnt[i];
  std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
  std::vector<std::pair<int32_t, float>> tmp(static_cast<size_t>(nnz));
  for (int64_t i = 0; i < nnz; ++i) tmp[pos[r[i]]++] = {static_cast<int32_t>(c[i]), v[i]};
  out.rowptr.assign(static_cast<size_t>(n_rows) + 1, 0);
  out.col.reserve(nnz);
  out.val.reserve(nnz);
  for (int32_t row = 0; row < n_rows; ++row) {
    auto b = tmp.begin() + cnt[row], e = tmp.begin() + cnt[row + 1];
    std::stable_sort(b, e, [](const auto& x, const auto& y) { return x.first < y.first; });
    for (auto it = b; it != e; ++it) {
      if (!out.col.empty() && static_cast<int64_t>(out.col.size()) > out.rowptr[row] &&
          out.col.back() == it->first) {
        out.val.back() += it->second;  // coalesce duplicates
      } else {
        out.col.push_back(it->first);
        out.val.push_back(it->second);
      }
    }
    out.rowptr[row + 1] = static_cast<int32_t>(out.col.size());
  }
  return out;
}

static HostCsr transpose(const HostCsr& a) {
  const int64_t nnz = static_cast<int64_t>(a.col.size());
  std::vector<int64_t> r(nnz), c(nnz);
  for (int32_t row = 0; row < a.n_rows; ++row)
    for (int32_t e = a.rowptr[row]; e < a.rowptr[row + 1]; ++e) {
      r[e] = a.col[e];
      c[e] = row;
    }
  return build_csr(a.n_cols, a.n_rows, nnz, r.data(), c.data(), a.val.data());
}

struct HostRb {
  int32_t R = 0, n_blocks = 0, max_union = 0;
  std::vector<int32_t> perm;  // original row of permuted position p (padded to n_blocks * R with -1); empty = natural order
  std::vector<int32_t> blkptr, ucol;
  std::vector<float> uval;
  // tiles of DSW_TILE_BLOCKS row-blocks
  int32_t n_tiles = 0, tile_rows_max = 0;
  std::vector<int32_t> tile_ptr, tile_row;
  std::vector<uint16_t> lidx;
  int32_t tile_pieces_max = 0;
  std::vector<int32_t> tpc_ptr, tpc_row;
  std::vector<uint32_t> tpc_meta;
  int32_t tile_len_max = 0;
  std::vector<int32_t> tp_ptr;
  std::vector<float> tp_val;      // 4 floats per (step, slot)
  std::vector<uint32_t> tp_off;
  int32_t tile_deps_max = 0;
  std::vector<int32_t> tdep_ptr, tdep_idx;
  // Two-phase staging of the chain kernel (0 = none): the pieces of the run of source rows that holds the tile's own rows come
  // first in the tile's piece list ("group A": half of the gathered rows, 2-3 boxes out of ~22), every row-block's panel
  // lists the entries that gather from group A first, and the first `steps_a[w]` entry steps of compute warp w touch nothing
  // else — the entry loop starts on them while the other boxes are still being issued / in flight.
  // packed: pieces of A | rows of A << 8 | steps_a[0] << 16 | steps_a[1] << 24
  std::vector<uint32_t> tile_split;
};

static void build_tiles(HostRb& rb) {
  rb.n_tiles = (rb.n_blocks + DSW_TILE_BLOCKS - 1) / DSW_TILE_BLOCKS;
  rb.tile_ptr.assign(static_cast<size_t>(rb.n_tiles) + 1, 0);
  rb.lidx.assign(rb.ucol.size(), 0);
  std::vector<int32_t> rows;
  bool fits16 = true;
  for (int32_t t = 0; t < rb.n_tiles; ++t) {
    const int32_t e0 = rb.blkptr[t * DSW_TILE_BLOCKS];
    const int32_t e1 = rb.blkptr[std::min(rb.n_blocks, (t + 1) * DSW_TILE_BLOCKS)];
    rows.assign(rb.ucol.begin() + e0, rb.ucol.begin() + e1);
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    if (rows.size() > 65535) fits16 = false;
    for (int32_t e = e0; e < e1 && fits16; ++e)
      rb.lidx[e] = static_cast<uint16_t>(std::lower_bound(rows.begin(), rows.end(), rb.ucol[e]) - rows.begin());
    rb.tile_row.insert(rb.tile_row.end(), rows.begin(), rows.end());
    rb.tile_ptr[t + 1] = static_cast<int32_t>(rb.tile_row.size());
    rb.tile_rows_max = std::max<int32_t>(rb.tile_rows_max, static_cast<int32_t>(rows.size()));
  }
  if (!fits16) {  // tile layout unusable; kernels fall back
    rb.n_tiles = 0, rb.tile_rows_max = 0;
    return;
  }
  // TMA pieces: runs of consecutive source rows, cut into power-of-two boxes (<= 128 rows)
  rb.tpc_ptr.assign(static_cast<size_t>(rb.n_tiles) + 1, 0);
  for (int32_t t = 0; t < rb.n_tiles; ++t) {
    const int32_t r0 = rb.tile_ptr[t], r1 = rb.tile_ptr[t + 1];
    int32_t i = r0;
    while (i < r1) {
      int32_t j = i + 1;
      while (j < r1 && rb.tile_row[j] == rb.tile_row[j - 1] + 1) ++j;
      int32_t len = j - i, at = i;
      while (len > 0) {
        int32_t lg = 0;
        while ((2 << lg) <= len && lg < 7) ++lg;
        rb.tpc_row.push_back(rb.tile_row[at]);
        rb.tpc_meta.push_back((static_cast<uint32_t>(at - r0) << 8) | static_cast<uint32_t>(lg));
        at += 1 << lg;
        len -= 1 << lg;
      }
      i = j;
    }
    rb.tpc_ptr[t + 1] = static_cast<int32_t>(rb.tpc_row.size());
    rb.tile_pieces_max = std::max(rb.tile_pieces_max, rb.tpc_ptr[t + 1] - rb.tpc_ptr[t]);
  }
  // Group A of every tile (see HostRb::tile_split): local row range [a_lo, a_hi] of the run that holds the tile's own first
  // row, its pieces moved to the front of the tile's piece list.  (Natural order only; a tile without such a run has no A.)
  std::vector<int32_t> a_lo(static_cast<size_t>(rb.n_tiles), 0), a_hi(static_cast<size_t>(rb.n_tiles), -1);
  std::vector<int32_t> a_pieces(static_cast<size_t>(rb.n_tiles), 0);
  if (rb.perm.empty()) {
    std::vector<std::pair<int32_t, uint32_t>> front, back;
    for (int32_t t = 0; t < rb.n_tiles; ++t) {
      const int32_t r0 = rb.tile_ptr[t], r1 = rb.tile_ptr[t + 1];
      const int32_t own = t * DSW_TILE_BLOCKS * rb.R;
      const auto it = std::lower_bound(rb.tile_row.begin() + r0, rb.tile_row.begin() + r1, own);
      if (it == rb.tile_row.begin() + r1 || *it != own) continue;
      int32_t lo = static_cast<int32_t>(it - rb.tile_row.begin()), hi = lo;
      while (lo > r0 && rb.tile_row[lo - 1] + 1 == rb.tile_row[lo]) --lo;
      while (hi + 1 < r1 && rb.tile_row[hi + 1] == rb.tile_row[hi] + 1) ++hi;
      if (hi - lo + 1 > 255 || hi - lo + 1 == r1 - r0) continue;  // (everything in one run: nothing to overlap)
      front.clear(), back.clear();
      for (int32_t i = rb.tpc_ptr[t]; i < rb.tpc_ptr[t + 1]; ++i) {
        const int32_t first = static_cast<int32_t>(rb.tpc_meta[i] >> 8) + r0;
        (first >= lo && first <= hi ? front : back).emplace_back(rb.tpc_row[i], rb.tpc_meta[i]);
      }
      if (front.empty() || front.size() > 255) continue;
      int32_t i = rb.tpc_ptr[t];
      for (const auto& pc : front) rb.tpc_row[i] = pc.first, rb.tpc_meta[i] = pc.second, ++i;
      for (const auto& pc : back) rb.tpc_row[i] = pc.first, rb.tpc_meta[i] = pc.second, ++i;
      a_lo[t] = lo - r0, a_hi[t] = hi - r0, a_pieces[t] = static_cast<int32_t>(front.size());
    }
  }
  // entry-major padded panels
  rb.tp_ptr.assign(static_cast<size_t>(rb.n_tiles) + 1, 0);
  rb.tile_split.assign(static_cast<size_t>(rb.n_tiles), 0u);
  std::vector<int32_t> ord;
  for (int32_t t = 0; t < rb.n_tiles; ++t) {
    const int32_t b0 = t * DSW_TILE_BLOCKS, b1 = std::min(rb.n_blocks, b0 + DSW_TILE_BLOCKS);
    int32_t len = 0;
    for (int32_t b = b0; b < b1; ++b) len = std::max(len, rb.blkptr[b + 1] - rb.blkptr[b]);
    rb.tile_len_max = std::max(rb.tile_len_max, len);
    const size_t base = static_cast<size_t>(rb.tp_ptr[t]) * DSW_TILE_BLOCKS;
    const bool split = a_hi[t] >= a_lo[t];
    // the tile's panels end with DSW_PANEL_PAD all-zero steps, so that one bulk copy stages them ready to use; padding steps
    // (zero weights) read a row of group A, which has always landed when a step runs
    rb.tp_val.resize((base + static_cast<size_t>(len + DSW_PANEL_PAD) * DSW_TILE_BLOCKS) * 4, 0.f);
    rb.tp_off.resize(base + static_cast<size_t>(len + DSW_PANEL_PAD) * DSW_TILE_BLOCKS, split ? static_cast<uint32_t>(a_lo[t]) * 256u : 0u);
    int32_t steps_a[2] = {INT32_MAX, INT32_MAX}, steps[2] = {0, 0};
    for (int32_t b = b0; b < b1; ++b) {
      // entries that gather from group A first (stable: ascending columns inside each part)
      ord.clear();
      for (int32_t e = rb.blkptr[b]; e < rb.blkptr[b + 1]; ++e)
        if (split && rb.lidx[e] >= a_lo[t] && rb.lidx[e] <= a_hi[t]) ord.push_back(e);
      const int32_t n_a = static_cast<int32_t>(ord.size());
      for (int32_t e = rb.blkptr[b]; e < rb.blkptr[b + 1]; ++e)
        if (!(split && rb.lidx[e] >= a_lo[t] && rb.lidx[e] <= a_hi[t])) ord.push_back(e);
      const int32_t n_all = static_cast<int32_t>(ord.size());
      const int w = (b - b0) / (DSW_TILE_BLOCKS / 2);
      steps[w] = std::max(steps[w], n_all);
      if (n_a < n_all) steps_a[w] = std::min(steps_a[w], n_a);  // (a row-block gathering from A only constrains nothing)
      for (int32_t u = 0; u < n_all; ++u) {
        const int32_t e = ord[u];
        const size_t at = base + static_cast<size_t>(u) * DSW_TILE_BLOCKS + (b - b0);
        for (int r = 0; r < 4; ++r) rb.tp_val[at * 4 + r] = rb.uval[static_cast<size_t>(e) * 4 + r];
        rb.tp_off[at] = static_cast<uint32_t>(rb.lidx[e]) * 256u;
      }
    }
    if (split) {
      uint32_t sa[2];
      for (int w = 0; w < 2; ++w) sa[w] = static_cast<uint32_t>(std::min({steps_a[w], (steps[w] + 1) & ~1, 254})) & ~1u;
      // safety net for the invariant the kernel relies on (a step of the first phase must not touch a row that lands on the
      // second barrier): re-read the panel just written; a tile that violates it is staged in one phase
      bool holds = true;
      for (int32_t sl = 0; sl < b1 - b0 && holds; ++sl)
        for (uint32_t u = 0; u < sa[sl / (DSW_TILE_BLOCKS / 2)] && holds; ++u) {
          const int32_t local = static_cast<int32_t>(rb.tp_off[base + static_cast<size_t>(u) * DSW_TILE_BLOCKS + sl] / 256u);
          holds = local >= a_lo[t] && local <= a_hi[t];
        }
      if (holds)
        rb.tile_split[t] = static_cast<uint32_t>(a_pieces[t]) | static_cast<uint32_t>(a_hi[t] - a_lo[t] + 1) << 8 | sa[0] << 16 | sa[1] << 24;
    }
    rb.tp_ptr[t + 1] = rb.tp_ptr[t] + len + DSW_PANEL_PAD;
  }
  // Tile dependencies of the fused chain kernel: tile t of hop k may start once hop k-1 has finished every tile
  // that holds one of t's source rows (square operators in natural order only; own tile always included).
  rb.tdep_ptr.assign(static_cast<size_t>(rb.n_tiles) + 1, 0);
  if (rb.perm.empty()) {
    const int32_t rows_per_tile = 4 * DSW_TILE_BLOCKS;
    std::vector<int32_t> deps;
    for (int32_t t = 0; t < rb.n_tiles; ++t) {
      deps.clear();
      deps.push_back(t);
      for (int32_t i = rb.tile_ptr[t]; i < rb.tile_ptr[t + 1]; ++i) deps.push_back(rb.tile_row[i] / rows_per_tile);
      std::sort(deps.begin(), deps.end());
      deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
      rb.tdep_idx.insert(rb.tdep_idx.end(), deps.begin(), deps.end());
      rb.tdep_ptr[t + 1] = static_cast<int32_t>(rb.tdep_idx.size());
      rb.tile_deps_max = std::max<int32_t>(rb.tile_deps_max, static_cast<int32_t>(deps.size()));
    }
  }
}

// `order` (optional): permuted position -> original row; row-block b then holds rows order[b*R .. b*R+R-1].
static HostRb build_rb(const HostCsr& a, int32_t R, const std::vector<int32_t>* order = nullptr) {
  HostRb rb;
  rb.R = R;
  rb.n_blocks = (a.n_rows + R - 1) / R;
  if (order) {
    rb.perm = *order;
    rb.perm.resize(static_cast<size_t>(rb.n_blocks) * R, -1);
  }
  rb.blkptr.assign(static_cast<size_t>(rb.n_blocks) + 1, 0);
  std::vector<int32_t> uni;
  int32_t rows[8];
  for (int32_t blk = 0; blk < rb.n_blocks; ++blk) {
    for (int32_t r = 0; r < R; ++r) {
      const int32_t pos = blk * R + r;
      rows[r] = order ? rb.perm[pos] : (pos < a.n_rows ? pos : -1);
    }
    uni.clear();
    for (int32_t r = 0; r < R; ++r)
      if (rows[r] >= 0) uni.insert(uni.end(), a.col.begin() + a.rowptr[rows[r]], a.col.begin() + a.rowptr[rows[r] + 1]);
    std::sort(uni.begin(), uni.end());
    uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
    const size_t base = rb.ucol.size();
    rb.ucol.insert(rb.ucol.end(), uni.begin(), uni.end());
    rb.uval.resize((base + uni.size()) * R, 0.f);
    for (int32_t r = 0; r < R; ++r) {
      if (rows[r] < 0) continue;
      for (int32_t e = a.rowptr[rows[r]]; e < a.rowptr[rows[r] + 1]; ++e) {
        const size_t u = std::lower_bound(uni.begin(), uni.end(), a.col[e]) - uni.begin();
        rb.uval[(base + u) * R + r] = a.val[e];
      }
    }
    rb.blkptr[blk + 1] = static_cast<int32_t>(rb.ucol.size());
    rb.max_union = std::max<int32_t>(rb.max_union, static_cast<int32_t>(uni.size()));
  }
  if (R == 4) build_tiles(rb);
  return rb;
}

// Locality-preserving row order for operators whose natural order has none (e.g. a row-major lat-lon grid:
// 64 consecutive nodes are a strip, and a tile of the hop kernel would gather ~2.5x more source rows than
// a compact patch).  Greedy region growing over the operator's own graph: clusters of one tile's worth of rows
// are grown breadth-first from a seed; the next seed is the oldest still-unvisited row on the frontier of what
// has been clustered so far, so that consecutive clusters are adjacent too.  Square operators only.
static std::vector<int32_t> locality_order(const HostCsr& a, int32_t cluster_rows) {
  const int32_t n = a.n_rows;
  std::vector<int32_t> order;
  order.reserve(n);
  std::vector<uint8_t> state(n, 0);  // 0 = untouched, 1 = on the global frontier, 2 = clustered
  std::vector<int32_t> frontier, queue;
  std::vector<int32_t> stamp(a.n_cols, 0);
  int32_t stamp_id = 0;
  size_t frontier_head = 0;
  int32_t next_natural = 0;
  while (static_cast<int32_t>(order.size()) < n) {
    int32_t seed = -1;
    while (frontier_head < frontier.size()) {
      const int32_t c = frontier[frontier_head++];
      if (state[c] != 2) { seed = c; break; }
    }
    if (seed < 0) {
      while (state[next_natural] == 2) ++next_natural;
      seed = next_natural;
    }
    queue.clear();
    queue.push_back(seed);
    state[seed] = 2;
    for (size_t head = 0; head < queue.size(); ++head) {
      const int32_t r = queue[head];
      for (int32_t e = a.rowptr[r]; e < a.rowptr[r + 1]; ++e) {
        const int32_t c = a.col[e];
        if (c < n && state[c] != 2 && static_cast<int32_t>(queue.size()) < cluster_rows) {
          state[c] = 2;
          queue.push_back(c);
        } else if (c < n && state[c] == 0) {
          state[c] = 1;
          frontier.push_back(c);
        }
      }
    }
    // Inside the cluster, group the rows four by four (one row-block) so that a group's rows share as many
    // columns as possible: start from the oldest ungrouped row, then add three times the ungrouped row with
    // the largest overlap with the group's column union.
    std::vector<uint8_t> used(queue.size(), 0);
    for (size_t first = 0; first < queue.size(); ++first) {
      if (used[first]) continue;
      used[first] = 1;
      order.push_back(queue[first]);
      ++stamp_id;
      for (int32_t e = a.rowptr[queue[first]]; e < a.rowptr[queue[first] + 1]; ++e) stamp[a.col[e]] = stamp_id;
      for (int pick = 0; pick < 3; ++pick) {
        int best = -1, best_overlap = -1;
        for (size_t i = first + 1; i < queue.size(); ++i) {
          if (used[i]) continue;
          int overlap = 0;
          for (int32_t e = a.rowptr[queue[i]]; e < a.rowptr[queue[i] + 1]; ++e) overlap += stamp[a.col[e]] == stamp_id;
          if (overlap > best_overlap) best_overlap = overlap, best = static_cast<int>(i);
        }
        if (best < 0) break;
        used[best] = 1;
        order.push_back(queue[best]);
        for (int32_t e = a.rowptr[queue[best]]; e < a.rowptr[queue[best] + 1]; ++e) stamp[a.col[e]] = stamp_id;
      }
    }
  }
  return order;
}

// Row-major 2-D grids (equiangular lat x lon sampling in lat-major order, cfg5): 64 consecutive nodes are a strip along
// one latitude row, so a tile gathers its strip plus the strips of the rows above and below (~280 source rows where a
// compact patch needs ~150).  The grid width W shows in the operator itself: the column offsets +-1 (same row) and +-W
// (the rows above / below) occur in most rows.  Returns W, or 0 when the operator does not look like such a grid.
static int32_t detect_row_major_width(const HostCsr& a) {
  const int32_t n = a.n_rows;
  if (n != a.n_cols || n < 256) return 0;
  const int32_t limit = std::min<int32_t>(n / 4, 1 << 16);
  std::vector<int32_t> cnt(static_cast<size_t>(limit) + 1, 0);
  for (int32_t r = 0; r < n; ++r)
    for (int32_t e = a.rowptr[r]; e < a.rowptr[r + 1]; ++e) {
      const int32_t d = a.col[e] - r;
      if (d > 0 && d <= limit) cnt[d]++;
    }
  if (cnt[1] < n / 2) return 0;
  for (int32_t w = 8; w <= limit; ++w)
    if (cnt[w] >= n / 2 && n % w == 0) return w;
  return 0;
}

// Patch order of a row-major H x W grid: tiles of th x tw nodes (one tile of the hop kernel), inside a tile 2 x 2 blocks
// (one row-block of four nodes with a small column union), inside a block row-major.
static std::vector<int32_t> grid_patch_order(int32_t H, int32_t W, int32_t th, int32_t tw) {
  std::vector<int32_t> order;
  order.reserve(static_cast<size_t>(H) * W);
  for (int32_t ty = 0; ty < H; ty += th)
    for (int32_t tx = 0; tx < W; tx += tw)
      for (int32_t by = 0; by < th; by += 2)
        for (int32_t bx = 0; bx < tw; bx += 2)
          for (int32_t dy = 0; dy < 2; ++dy)
            for (int32_t dx = 0; dx < 2; ++dx) {
              const int32_t y = ty + by + dy, x = tx + bx + dx;
              if (y < H && x < W) order.push_back(y * W + x);
            }
  return order;
}

static double avg_tile_rows(const HostRb& rb) {
  return rb.n_tiles > 0 ? static_cast<double>(rb.tile_row.size()) / rb.n_tiles : 0.0;
}

// Row-block layout in the natural order, or — when that order gathers far more source rows per tile than a
// compact patch would and the region-grown order is clearly better — in the locality order.
static HostRb build_rb_auto(const HostCsr& a, int32_t R) {
  HostRb nat = build_rb(a, R);
  const int64_t opt = g_options[DSW_OPT_PLAN_PERMUTE].load(std::memory_order_relaxed);  // 0 auto, 1 never, 2 always
  if (R != 4 || a.n_rows != a.n_cols || nat.n_tiles < 4 || opt == 1) return nat;
  const double rows_per_tile = 4.0 * DSW_TILE_BLOCKS;
  // Row-major lat-lon grids: tile 4 x 16 patches instead of 1 x 64 strips (cfg5: 281 -> 151 gathered rows per tile, as
  // many as a HEALPix nested tile; 16 instead of 10 TMA boxes).  Tried whenever the natural order gathers >= 3x its rows.
  if (opt != 2 && avg_tile_rows(nat) >= 3.0 * rows_per_tile) {
    const int32_t W = detect_row_major_width(a);
    if (W > 0) {
      const std::vector<int32_t> order = grid_patch_order(a.n_rows / W, W, 4, 16);
      HostRb grid = build_rb(a, R, &order);
      if (grid.n_tiles > 0 && avg_tile_rows(grid) < 0.75 * avg_tile_rows(nat)) return grid;
    }
  }
  // Automatic mode permutes only when the natural order is hopeless (a tile would gather > 6x its own rows, e.g.
  // randomly numbered nodes: the tile kernel would not even fit two teams).  A row-major lat-lon grid (4.4x,
  // cfg5) measured 8 % *slower* permuted: the region-grown patches need 3-4x more, smaller TMA boxes.
  if (opt != 2 && avg_tile_rows(nat) < 6.0 * rows_per_tile) return nat;
  const std::vector<int32_t> order = locality_order(a, 4 * DSW_TILE_BLOCKS);
  HostRb loc = build_rb(a, R, &order);
  if (loc.n_tiles == 0) return nat;
  if (opt != 2 && avg_tile_rows(loc) > 0.85 * avg_tile_rows(nat)) return nat;
  return loc;
}

// Cost model (cycles per row-block per 64 features, one SM): gathers cost 2 cycles of the 128 B/clk
// L1 path per union entry, FMAs cost R/2 cycles per union entry at 128 FMA/clk.
static int32_t pick_rb_rows(const HostCsr& a) {
  const int64_t nnz = static_cast<int64_t>(a.col.size());
  if (a.n_rows < 8 || nnz == 0) return 0;
  double best = 2.0 * nnz;  // plain CSR: L1-bound
  int32_t best_r = 0;
  for (int32_t R : {2, 4}) {
    int64_t uni_total = 0;
    std::vector<int32_t> uni;
    for (int32_t r0 = 0; r0 < a.n_rows; r0 += R) {
      const int32_t r1 = std::min(a.n_rows, r0 + R);
      uni.assign(a.col.begin() + a.rowptr[r0], a.col.begin() + a.rowptr[r1]);
      std::sort(uni.begin(), uni.end());
      uni_total += std::unique(uni.begin(), uni.end()) - uni.begin();
    }
    const double cost = uni_total * std::max(2.0, 0.5 * R);
    if (cost < 0.85 * best) {
      best = cost;
      best_r = R;
    }
  }
  return best_r;
}

template <class T>
static cudaError_t upload(T** dst, const std::vector<T>& src, cudaStream_t st) {
  *dst = nullptr;
  const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dst), bytes);
  if (e != cudaSuccess) return e;
  if (!src.empty()) e = cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st);
  return e;
}

static cudaError_t upload_csr(dsw_csr* d, const HostCsr& h, cudaStream_t st) {
  d->n_rows = h.n_rows;
  d->n_cols = h.n_cols;
  d->nnz = static_cast<int64_t>(h.col.size());
  d->max_row_nnz = 0;
  for (int32_t r = 0; r < h.n_rows; ++r) d->max_row_nnz = std::max(d->max_row_nnz, h.rowptr[r + 1] - h.rowptr[r]);
  cudaError_t e;
  if ((e = upload(&d->rowptr, h.rowptr, st)) != cudaSuccess) return e;
  if ((e = upload(&d->col, h.col, st)) != cudaSuccess) return e;
  return upload(&d->val, h.val, st);
}

static cudaError_t upload_rb(dsw_rb* d, const HostRb& h, cudaStream_t st, bool square) {
  d->R = h.R;
  d->n_blocks = h.n_blocks;
  d->max_union = h.max_union;
  d->total_union = static_cast<int64_t>(h.ucol.size());
  d->tile_entries_max = 0;
  for (int32_t b0 = 0; b0 < h.n_blocks; b0 += 32)
    d->tile_entries_max = std::max(d->tile_entries_max, h.blkptr[std::min(h.n_blocks, b0 + 32)] - h.blkptr[b0]);
  cudaError_t e;
  if ((e = upload(&d->blkptr, h.blkptr, st)) != cudaSuccess) return e;
  if ((e = upload(&d->ucol, h.ucol, st)) != cudaSuccess) return e;
  if ((e = upload(&d->uval, h.uval, st)) != cudaSuccess) return e;
  if (!h.perm.empty() && (e = upload(&d->perm, h.perm, st)) != cudaSuccess) return e;
  d->n_tiles = h.n_tiles;
  d->tile_rows_max = h.tile_rows_max;
  if (h.n_tiles > 0) {
    if ((e = upload(&d->tile_ptr, h.tile_ptr, st)) != cudaSuccess) return e;
    if ((e = upload(&d->tile_row, h.tile_row, st)) != cudaSuccess) return e;
    if ((e = upload(&d->lidx, h.lidx, st)) != cudaSuccess) return e;
    d->tile_pieces_max = h.tile_pieces_max;
    if ((e = upload(&d->tpc_ptr, h.tpc_ptr, st)) != cudaSuccess) return e;
    if ((e = upload(&d->tpc_row, h.tpc_row, st)) != cudaSuccess) return e;
    if ((e = upload(&d->tpc_meta, h.tpc_meta, st)) != cudaSuccess) return e;
    d->tile_len_max = h.tile_len_max;
    if ((e = upload(&d->tp_ptr, h.tp_ptr, st)) != cudaSuccess) return e;
    float* tv = nullptr;
    if ((e = upload(&tv, h.tp_val, st)) != cudaSuccess) return e;
    d->tp_val = reinterpret_cast<float4*>(tv);
    if ((e = upload(&d->tp_off, h.tp_off, st)) != cudaSuccess) return e;
    const size_t cnt_bytes = (size_t)HOP_CNT_SLOTS * (h.n_tiles + 1) * sizeof(int32_t);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&d->hop_cnt), cnt_bytes)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(d->hop_cnt, 0, cnt_bytes, st)) != cudaSuccess) return e;
    d->hop_ring = new std::atomic<uint32_t>(0);
    // fused chain kernel: dependency lists + sync words (square operators whose tiles gather from few tiles)
    if (h.tile_deps_max > 0 && h.tile_deps_max <= DSW_CHAIN_MAX_DEPS && square) {
      if ((e = upload(&d->tdep_ptr, h.tdep_ptr, st)) != cudaSuccess) return e;
      if ((e = upload(&d->tdep_idx, h.tdep_idx, st)) != cudaSuccess) return e;
      d->tile_deps_max = h.tile_deps_max;
      {
        std::vector<int4> meta(static_cast<size_t>(h.n_tiles) * 2);
        std::vector<int2> pcs(static_cast<size_t>(h.n_tiles) * h.tile_pieces_max, make_int2(-1, 0));
        std::vector<int32_t> deps(static_cast<size_t>(h.n_tiles) * h.tile_deps_max, -1);
        for (int32_t t = 0; t < h.n_tiles; ++t) {
          const int32_t np = h.tpc_ptr[t + 1] - h.tpc_ptr[t], nd = h.tdep_ptr[t + 1] - h.tdep_ptr[t];
          meta[2 * t] = make_int4(h.tp_ptr[t], h.tp_ptr[t + 1] - h.tp_ptr[t], h.tile_ptr[t + 1] - h.tile_ptr[t], np);
          // entry-loop trip counts of the tile's two compute warps (row-block slots 0-7 / 8-15): longest union, rounded up
          // to the two-step software pipeline — handed to the team through the item descriptor
          int32_t wl[2] = {0, 0};
          for (int32_t sl = 0; sl < DSW_TILE_BLOCKS; ++sl) {
            const int32_t blk = t * DSW_TILE_BLOCKS + sl;
            if (blk < h.n_blocks) wl[sl / (DSW_TILE_BLOCKS / 2)] = std::max(wl[sl / (DSW_TILE_BLOCKS / 2)], h.blkptr[blk + 1] - h.blkptr[blk]);
          }
          meta[2 * t + 1] = make_int4(nd, (wl[0] + 1) & ~1, (wl[1] + 1) & ~1, static_cast<int32_t>(h.tile_split[t]));
          for (int32_t i = 0; i < np; ++i)
            pcs[static_cast<size_t>(t) * h.tile_pieces_max + i] =
                make_int2(static_cast<int32_t>(h.tpc_meta[h.tpc_ptr[t] + i]), h.tpc_row[h.tpc_ptr[t] + i]);
          for (int32_t i = 0; i < nd; ++i) deps[static_cast<size_t>(t) * h.tile_deps_max + i] = h.tdep_idx[h.tdep_ptr[t] + i];
        }
        if ((e = upload(&d->tile_meta, meta, st)) != cudaSuccess) return e;
        if ((e = upload(&d->tpc_fix, pcs, st)) != cudaSuccess) return e;
        if ((e = upload(&d->tdep_fix, deps, st)) != cudaSuccess) return e;
      }
      d->chain_flag_cap = std::max(64 * h.n_tiles, 32768);
      const size_t words = (size_t)DSW_CHAIN_SETS * (DSW_CHAIN_HDR + (size_t)d->chain_flag_cap);
      if ((e = cudaMalloc(reinterpret_cast<void**>(&d->chain_sync), words * sizeof(int32_t))) != cudaSuccess) return e;
      if ((e = cudaMemsetAsync(d->chain_sync, 0, words * sizeof(int32_t), st)) != cudaSuccess) return e;
      d->chain_ring = new std::atomic<uint32_t>(0);
    }
  }
  return cudaSuccess;
}

static void free_csr(dsw_csr* c) {
  cudaFree(c->rowptr);
  cudaFree(c->col);
  cudaFree(c->val);
  *c = dsw_csr{};
}
static void free_rb(dsw_rb* r) {
  cudaFree(r->blkptr);
  cudaFree(r->ucol);
  cudaFree(r->uval);
  cudaFree(r->tile_ptr);
  cudaFree(r->tile_row);
  cudaFree(r->lidx);
  cudaFree(r->tpc_ptr);
  cudaFree(r->tpc_row);
  cudaFree(r->tpc_meta);
  cudaFree(r->tp_ptr);
  cudaFree(r->tp_val);
  cudaFree(r->tp_off);
  cudaFree(r->perm);
  cudaFree(r->hop_cnt);
  delete r->hop_ring;
  cudaFree(r->tdep_ptr);
  cudaFree(r->tdep_idx);
  cudaFree(r->tile_meta);
  cudaFree(r->tpc_fix);
  cudaFree(r->tdep_fix);
  cudaFree(r->chain_sync);
  delete r->chain_ring;
  *r = dsw_rb{};
}

}  // namespace dsw

using namespace dsw;

extern "C" {

int dsw_version(void) { return DSW_VERSION; }

const char* dsw_strerror(int s) {
  switch (s) {
    case DSW_OK: return "success";
    case DSW_ERR_BAD_ARGUMENT: return "bad argument (null pointer, non-positive size or bad enum)";
    case DSW_ERR_SHAPE: return "shape mismatch between plan and tensors";
    case DSW_ERR_WORKSPACE: return "workspace missing or too small";
    case DSW_ERR_UNSUPPORTED: return "unsupported configuration";
    case DSW_ERR_CUDA: return "CUDA runtime error (see dsw_last_cuda_error_string)";
    case DSW_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required)";
    case DSW_ERR_ALIGNMENT: return "pointer or stride alignment violated";
    default: return "unknown dsw status";
  }
}

int dsw_last_cuda_error(void) { return static_cast<int>(last_cuda_error()); }
const char* dsw_last_cuda_error_string(void) { return cudaGetErrorString(last_cuda_error()); }

int dsw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

int64_t dsw_launch_count(void) { return g_launches.load(); }
int dsw_set_mix_mode(int mode) {
  if (mode < 0 || mode > 1) return DSW_ERR_BAD_ARGUMENT;
  g_mix_mode.store(mode);
  return DSW_OK;
}
int dsw_get_mix_mode(void) { return g_mix_mode.load(); }
int dsw_set_option(int key, int64_t value) {
  if (key < 0 || key >= DSW_OPT_COUNT || value < 0) return DSW_ERR_BAD_ARGUMENT;
  g_options[key].store(value);
  return DSW_OK;
}
int64_t dsw_get_option(int key) {
  if (key < 0 || key >= DSW_OPT_COUNT) return -1;
  return g_options[key].load();
}

int dsw_plan_create(int32_t n_rows, int32_t n_cols, int64_t nnz, const int64_t* coo_row,
                    const int64_t* coo_col, const float* coo_val, void* stream, dsw_plan** out) {
  if (!out) return DSW_ERR_BAD_ARGUMENT;
  *out = nullptr;
  if (n_rows <= 0 || n_cols <= 0 || nnz < 0) return DSW_ERR_BAD_ARGUMENT;
  if (nnz > 0 && (!coo_row || !coo_col || !coo_val)) return DSW_ERR_BAD_ARGUMENT;
  if (nnz > INT32_MAX) return DSW_ERR_UNSUPPORTED;
  if (dsw_device_count() <= 0) return DSW_ERR_NO_DEVICE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  std::vector<int64_t> r(nnz), c(nnz);
  std::vector<float> v(nnz);
  if (nnz > 0) {
    DSW_CUDA_TRY(cudaStreamSynchronize(st));
    DSW_CUDA_TRY(cudaMemcpy(r.data(), coo_row, nnz * sizeof(int64_t), cudaMemcpyDefault));
    DSW_CUDA_TRY(cudaMemcpy(c.data(), coo_col, nnz * sizeof(int64_t), cudaMemcpyDefault));
    DSW_CUDA_TRY(cudaMemcpy(v.data(), coo_val, nnz * sizeof(float), cudaMemcpyDefault));
  }
  for (int64_t i = 0; i < nnz; ++i)
    if (r[i] < 0 || r[i] >= n_rows || c[i] < 0 || c[i] >= n_cols) return DSW_ERR_SHAPE;

  HostCsr fwd = build_csr(n_rows, n_cols, nnz, r.data(), c.data(), v.data());
  HostCsr tr = transpose(fwd);

  dsw_plan* p = new dsw_plan();
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = upload_csr(&p->fwd, fwd, st);
  if (e == cudaSuccess) e = upload_csr(&p->tr, tr, st);
  if (e == cudaSuccess) {
    const int32_t rf = pick_rb_rows(fwd);
    if (rf > 0) e = upload_rb(&p->fwd_rb, build_rb_auto(fwd, rf), st, n_rows == n_cols);
  }
  if (e == cudaSuccess) {
    const int32_t rt = pick_rb_rows(tr);
    if (rt > 0) e = upload_rb(&p->tr_rb, build_rb_auto(tr, rt), st, n_rows == n_cols);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // host vectors go out of scope
  if (e != cudaSuccess) {
    set_cuda_error(e);
    dsw_plan_destroy(p);
    return DSW_ERR_CUDA;
  }
  *out = p;
  return DSW_OK;
}

void dsw_plan_destroy(dsw_plan* p) {
  if (!p) return;
  free_csr(&p->fwd);
  free_csr(&p->tr);
  free_rb(&p->fwd_rb);
  free_rb(&p->tr_rb);
  delete p;
}

int dsw_plan_shape(const dsw_plan* p, int32_t* n_rows, int32_t* n_cols, int64_t* nnz, int32_t* max_row_nnz) {
  if (!p) return DSW_ERR_BAD_ARGUMENT;
  if (n_rows) *n_rows = p->fwd.n_rows;
  if (n_cols) *n_cols = p->fwd.n_cols;
  if (nnz) *nnz = p->fwd.nnz;
  if (max_row_nnz) *max_row_nnz = p->fwd.max_row_nnz;
  return DSW_OK;
}

int64_t dsw_plan_operand_bytes(const dsw_plan* p) {
  if (!p) return 0;
  return 8 * p->fwd.nnz + 4 * (static_cast<int64_t>(p->fwd.n_rows) + 1);
}

}  // extern "C"
