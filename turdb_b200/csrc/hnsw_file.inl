// hnsw_file.inl — reader of the reference's on-disk index (`.hnsw`) and uploader to the device layout.
// Host code only; included by turdb_cuda.cu (shares fail()/CUDA_TRY).  SURVEY.md §8(f) rank 1.
//
// What is read, with the reference ranges it follows (all under /root/reference):
//   file = pages of 16384 B (MmapStorage::page, src/storage/mmap.rs:201-211); page 0 starts with the 128 B
//   HnswFileHeader (src/hnsw/storage.rs:98-119); node pages carry the 16 B PageHeader (page_type 0x10,
//   src/storage/page.rs:86-127) + the 48 B HnswPageHeader (storage.rs:371-383), a slot directory of 4 B
//   entries from byte 64 (storage.rs:322-369, 485-503) and node records in HnswNode's wire format
//   (src/hnsw/mod.rs:333-421).  Which slots count as nodes follows rebuild_row_id_map (mod.rs:836-859) and
//   read_node (mod.rs:906-911): active slots whose bytes parse.
//
// Dense node ids are assigned in (page, slot) order == allocate_node order (mod.rs:883-904).
// A NodeId that is referenced (neighbour list, entry point) but does not resolve to a readable active slot
// — deleted (mark_deleted, mod.rs:937-948), never written, or damaged — becomes a TOMBSTONE: a dense id
// after the real nodes with no neighbours, row_id 0 and a +inf vector.  That reproduces what search() does
// with an unreadable node: distance INFINITY, no neighbours, row_id 0 in the result (mod.rs:1111-1127,
// 1159-1171).
//
// The slot directory stores only 13 offset bits (storage.rs:338-344) while records are allocated from the
// page end (offset up to 16383), and the reference both writes and reads through the truncated offset, so
// on a page with more than ~39 records later records overwrite earlier ones and the page header.  This
// reader reads exactly what the reference would (the truncated offset) and FLAGS every page on which two
// live records, or a record and the slot directory, overlap (`n_suspect_pages`); it does not try to repair.

#include <sys/stat.h>

#include <cmath>
#include <limits>
#include <unordered_map>

namespace {

constexpr size_t kPageSize = 16384;        // PAGE_SIZE, src/config/constants.rs
constexpr size_t kFileHeaderSize = 128;    // FILE_HEADER_SIZE
constexpr size_t kHnswPageHeader = 64;     // HNSW_PAGE_HEADER_SIZE, storage.rs:94
constexpr uint8_t kPageTypeHnswNode = 0x10;  // PageType::HnswNode, src/storage/page.rs:94
const uint8_t kHnswMagic[16] = {'T', 'u', 'r', 'D', 'B', ' ', 'H', 'N', 'S', 'W', 0, 0, 0, 0, 0, 0};  // file_manager.rs:123

inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

struct ParsedNode {
  uint64_t row_id = 0;
  uint8_t max_level = 0;
  std::vector<uint64_t> l0;                 // NodeId keys (page << 16 | slot), stored order
  std::vector<std::vector<uint64_t>> upper;  // [max_level]
};

inline uint64_t node_key(uint32_t page, uint16_t slot) { return ((uint64_t)page << 16) | slot; }

// HnswNode::read_from, mod.rs:361-421.  false == the Err(...) cases.
bool parse_node(const uint8_t* buf, size_t len, ParsedNode* out) {
  if (len < 10) return false;
  out->row_id = rd64(buf);
  out->max_level = buf[8];
  const uint8_t l0_count = buf[9];
  if (l0_count > TURDB_MAX_L0_NEIGHBORS) return false;
  size_t off = 10;
  out->l0.clear();
  for (uint32_t i = 0; i < l0_count; ++i) {
    if (off + 6 > len) return false;
    out->l0.push_back(node_key(rd32(buf + off), rd16(buf + off + 4)));
    off += 6;
  }
  out->upper.assign(out->max_level, {});
  for (uint32_t l = 0; l < out->max_level; ++l) {
    if (off >= len) return false;
    const uint32_t cnt = buf[off++];
    for (uint32_t i = 0; i < cnt; ++i) {
      if (off + 6 > len) return false;
      out->upper[l].push_back(node_key(rd32(buf + off), rd16(buf + off + 4)));
      off += 6;
    }
  }
  return true;
}

}  // namespace

struct turdb_cuda_hnsw_file {
  turdb_cuda_hnsw_file_info info{};
  std::vector<uint64_t> row_ids;   // [n_total]
  std::vector<uint32_t> node_page;  // dense id -> NodeId
  std::vector<uint16_t> node_slot;
  std::vector<uint8_t> levels, l0_cnt, up_cnt;
  std::vector<uint32_t> l0_adj, up_base, up_adj;
};

static int32_t hnsw_file_parse(const uint8_t* data, uint64_t len, turdb_cuda_hnsw_file** out) {
  if (!data || !out) return fail(TURDB_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  // HnswFileHeader::from_bytes, storage.rs:158-173
  if (len < kFileHeaderSize) return fail(TURDB_ERR_INVALID_ARGUMENT, "buffer too small for HnswFileHeader: %llu < 128", (unsigned long long)len);
  if (memcmp(data, kHnswMagic, 16) != 0) return fail(TURDB_ERR_INVALID_ARGUMENT, "invalid HNSW file: magic bytes mismatch");
  auto* f = new (std::nothrow) turdb_cuda_hnsw_file();
  if (!f) return fail(TURDB_ERR_OUT_OF_MEMORY, "host allocation failed");
  turdb_cuda_hnsw_file_info& I = f->info;
  I.index_id = rd64(data + 16);
  I.table_id = rd64(data + 24);
  I.dimensions = rd16(data + 32);
  I.m = rd16(data + 34);
  I.m0 = rd16(data + 36);
  I.ef_construction = rd16(data + 38);
  I.ef_search = rd16(data + 40);
  I.distance_fn = data[42] == 1 ? 1 : (data[42] == 2 ? 2 : 0);     // storage.rs:227-233: unknown -> L2
  I.quantization = data[43] == 1 ? 1 : (data[43] == 2 ? 2 : 0);    // storage.rs:235-241
  const uint32_t ep_page = rd32(data + 44);
  const uint16_t ep_slot = rd16(data + 48);
  I.header_max_level = data[50];
  I.header_node_count = rd64(data + 52);
  I.header_vector_count = rd64(data + 60);
  I.has_entry = ep_page != 0xFFFFFFFFu;  // HnswFileHeader::entry_point, storage.rs:243-252
  I.n_pages = (uint32_t)std::min<uint64_t>(len / kPageSize, 0xFFFFFFFFull);
  if (len % kPageSize) I.flags |= TURDB_HNSW_FILE_TRAILING_BYTES;

  std::unordered_map<uint64_t, uint32_t> dense;  // NodeId key -> dense id
  std::vector<ParsedNode> nodes;
  ParsedNode pn;
  for (uint32_t page = 1; page < I.n_pages; ++page) {
    const uint8_t* pg = data + (size_t)page * kPageSize;
    if (pg[0] != kPageTypeHnswNode) {  // HnswPageRef::from_bytes rejects it; rebuild_row_id_map skips the page
      I.n_foreign_pages += 1;
      continue;
    }
    const uint32_t slot_count = rd16(pg + 16);
    const size_t dir_end = kHnswPageHeader + (size_t)slot_count * 4;
    if (dir_end > kPageSize) {
      I.n_suspect_pages += 1;
      I.n_unreadable_slots += slot_count;
      continue;
    }
    std::vector<std::pair<uint32_t, uint32_t>> spans;  // live records [begin, end)
    bool suspect = false;
    for (uint32_t s = 0; s < slot_count; ++s) {
      // SlotEntry::decode, storage.rs:346-356
      const uint16_t os = rd16(pg + kHnswPageHeader + 4 * s);
      const uint32_t off = os & 0x1FFFu, status = (os >> 13) & 3u, size = rd16(pg + kHnswPageHeader + 4 * s + 2);
      if (status == 2) I.n_deleted_slots += 1;
      if (status != 1) continue;  // read_node_data: "slot is not active"
      if ((size_t)off + size > kPageSize || !parse_node(pg + off, size, &pn)) {
        I.n_unreadable_slots += 1;
        suspect = true;
        continue;
      }
      // add_neighbor_at_level never lets an upper list pass 16 entries (mod.rs:292-301); a longer one is
      // overwritten bytes, not a node: counted as unreadable (referrers get a tombstone)
      bool oversize = false;
      for (auto& lv : pn.upper) oversize |= lv.size() > TURDB_MAX_LEVEL_NEIGHBORS;
      if (oversize) {
        I.n_unreadable_slots += 1;
        suspect = true;
        continue;
      }
      if (off < dir_end) suspect = true;
      spans.emplace_back(off, off + size);
      if (nodes.size() >= 0x7FFFFFF0ull) {
        delete f;
        return fail(TURDB_ERR_UNSUPPORTED, "more than 2^31 nodes");
      }
      dense.emplace(node_key(page, (uint16_t)s), (uint32_t)nodes.size());
      f->node_page.push_back(page);
      f->node_slot.push_back((uint16_t)s);
      nodes.push_back(pn);
    }
    std::sort(spans.begin(), spans.end());
    for (size_t i = 1; i < spans.size(); ++i)
      if (spans[i].first < spans[i - 1].second) suspect = true;
    if (suspect) I.n_suspect_pages += 1;
  }
  I.n_nodes = nodes.size();

  // neighbour NodeIds -> dense ids; unresolved ones become tombstones
  auto resolve = [&](uint64_t key) -> uint32_t {
    auto it = dense.find(key);
    if (it != dense.end()) return it->second;
    const uint32_t id = (uint32_t)(I.n_nodes + I.n_tombstones);
    dense.emplace(key, id);
    f->node_page.push_back((uint32_t)(key >> 16));
    f->node_slot.push_back((uint16_t)(key & 0xFFFF));
    I.n_tombstones += 1;
    return id;
  };
  uint64_t n_slots = 0;
  for (const ParsedNode& nd : nodes) n_slots += nd.max_level;
  f->l0_adj.assign(nodes.size() * TURDB_MAX_L0_NEIGHBORS, TURDB_INVALID_NODE);
  f->up_adj.assign(n_slots * TURDB_MAX_LEVEL_NEIGHBORS, TURDB_INVALID_NODE);
  f->up_cnt.assign(n_slots, 0);
  uint64_t slot = 0;
  for (size_t i = 0; i < nodes.size(); ++i) {
    const ParsedNode& nd = nodes[i];
    f->row_ids.push_back(nd.row_id);
    f->levels.push_back(nd.max_level);
    f->l0_cnt.push_back((uint8_t)nd.l0.size());
    for (size_t j = 0; j < nd.l0.size(); ++j) f->l0_adj[i * TURDB_MAX_L0_NEIGHBORS + j] = resolve(nd.l0[j]);
    f->up_base.push_back(nd.max_level ? (uint32_t)slot : TURDB_INVALID_NODE);
    for (uint32_t l = 0; l < nd.max_level; ++l, ++slot) {
      f->up_cnt[slot] = (uint8_t)nd.upper[l].size();
      for (size_t j = 0; j < nd.upper[l].size(); ++j) f->up_adj[slot * TURDB_MAX_LEVEL_NEIGHBORS + j] = resolve(nd.upper[l][j]);
    }
  }
  I.entry = TURDB_INVALID_NODE;
  if (I.has_entry) I.entry = resolve(node_key(ep_page, ep_slot));
  // tombstones: level 0, no neighbours, row_id 0 (mod.rs:1163-1164)
  f->row_ids.resize(I.n_nodes + I.n_tombstones, 0);
  f->levels.resize(I.n_nodes + I.n_tombstones, 0);
  f->l0_cnt.resize(I.n_nodes + I.n_tombstones, 0);
  f->up_base.resize(I.n_nodes + I.n_tombstones, TURDB_INVALID_NODE);
  f->l0_adj.resize((I.n_nodes + I.n_tombstones) * TURDB_MAX_L0_NEIGHBORS, TURDB_INVALID_NODE);
  I.n_up_slots = n_slots;
  // greedy descent starts at header.max_level (mod.rs:1134); above the entry node's own level
  // neighbors_at_level returns [] (mod.rs:282-290) so those levels are no-ops: clamp
  I.max_level = I.header_max_level;
  if (I.entry != TURDB_INVALID_NODE && f->levels[I.entry] < I.max_level) {
    I.max_level = f->levels[I.entry];
    I.flags |= TURDB_HNSW_FILE_MAX_LEVEL_CLAMPED;
  }
  if (I.header_node_count != I.n_nodes) I.flags |= TURDB_HNSW_FILE_NODE_COUNT_MISMATCH;
  if (I.n_suspect_pages) I.flags |= TURDB_HNSW_FILE_SUSPECT_PAGES;
  if (I.n_tombstones) I.flags |= TURDB_HNSW_FILE_TOMBSTONES;
  *out = f;
  return TURDB_OK;
}

extern "C" {

// no exception crosses the ABI: the parser's containers may throw bad_alloc on a huge (or hostile) file
static int32_t hnsw_file_parse_noexcept(const uint8_t* bytes, uint64_t len, turdb_cuda_hnsw_file** out) {
  try {
    return hnsw_file_parse(bytes, len, out);
  } catch (const std::bad_alloc&) {
    if (out) *out = nullptr;
    return fail(TURDB_ERR_OUT_OF_MEMORY, "out of host memory while parsing a %llu-byte .hnsw file", (unsigned long long)len);
  } catch (...) {
    if (out) *out = nullptr;
    return fail(TURDB_ERR_INVALID_ARGUMENT, "unexpected failure while parsing the .hnsw file");
  }
}

int32_t turdb_cuda_hnsw_file_open_memory(const uint8_t* bytes, uint64_t len, turdb_cuda_hnsw_file** out) {
  return hnsw_file_parse_noexcept(bytes, len, out);
}

int32_t turdb_cuda_hnsw_file_open(const char* path, turdb_cuda_hnsw_file** out) {
  if (!path || !out) return fail(TURDB_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  FILE* fp = fopen(path, "rb");
  if (!fp) return fail(TURDB_ERR_INVALID_ARGUMENT, "failed to open '%s'", path);
  struct stat st {};
  if (fstat(fileno(fp), &st) != 0) {
    fclose(fp);
    return fail(TURDB_ERR_INVALID_ARGUMENT, "failed to stat '%s'", path);
  }
  std::vector<uint8_t> buf;
  try {
    buf.resize((size_t)st.st_size);
  } catch (...) {
    fclose(fp);
    return fail(TURDB_ERR_OUT_OF_MEMORY, "cannot buffer %lld bytes", (long long)st.st_size);
  }
  const size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), fp);
  fclose(fp);
  if (got != buf.size()) return fail(TURDB_ERR_INVALID_ARGUMENT, "short read on '%s'", path);
  return hnsw_file_parse_noexcept(buf.data(), buf.size(), out);
}

int32_t turdb_cuda_hnsw_file_close(turdb_cuda_hnsw_file* f) {
  delete f;
  return TURDB_OK;
}

int32_t turdb_cuda_hnsw_file_get_info(const turdb_cuda_hnsw_file* f, turdb_cuda_hnsw_file_info* out) {
  if (!f || !out) return fail(TURDB_ERR_INVALID_ARGUMENT, "null argument");
  *out = f->info;
  return TURDB_OK;
}

int32_t turdb_cuda_hnsw_file_nodes(const turdb_cuda_hnsw_file* f, uint64_t* out_row_ids, uint32_t* out_pages,
                                   uint16_t* out_slots) {
  if (!f) return fail(TURDB_ERR_INVALID_ARGUMENT, "file is null");
  const size_t n = f->row_ids.size();
  if (out_row_ids && n) memcpy(out_row_ids, f->row_ids.data(), n * 8);
  if (out_pages && n) memcpy(out_pages, f->node_page.data(), n * 4);
  if (out_slots && n) memcpy(out_slots, f->node_slot.data(), n * 2);
  return TURDB_OK;
}

int32_t turdb_cuda_hnsw_file_graph(const turdb_cuda_hnsw_file* f, const float* vectors, turdb_cuda_graph* out) {
  if (!f || !out) return fail(TURDB_ERR_INVALID_ARGUMENT, "null argument");
  const turdb_cuda_hnsw_file_info& I = f->info;
  memset(out, 0, sizeof(*out));
  out->dim = I.dimensions;
  out->max_level = I.max_level;
  out->n = I.n_nodes + I.n_tombstones;
  out->entry = I.entry;
  out->vectors = vectors;
  out->row_ids = f->row_ids.data();
  out->levels = f->levels.data();
  out->l0_adj = f->l0_adj.data();
  out->l0_cnt = f->l0_cnt.data();
  out->up_base = f->up_base.data();
  out->up_adj = f->up_adj.data();
  out->up_cnt = f->up_cnt.data();
  out->n_up_slots = I.n_up_slots;
  return TURDB_OK;
}

// vectors: [n_nodes][dim] in dense-id order (nullable when get_vector is given); present: nullable [n_nodes],
// 0 == the table has no vector for that row (get_vector -> None => distance INFINITY, mod.rs:1117-1120).
int32_t turdb_cuda_hnsw_file_upload(const turdb_cuda_hnsw_file* f, const float* vectors, const uint8_t* present,
                                    turdb_cuda_get_vector_fn get_vector, void* user, int32_t device,
                                    turdb_cuda_index** out) {
  if (!f || !out) return fail(TURDB_ERR_INVALID_ARGUMENT, "null argument");
  if (!vectors && !get_vector && f->info.n_nodes) return fail(TURDB_ERR_INVALID_ARGUMENT, "neither vectors nor get_vector given");
  const turdb_cuda_hnsw_file_info& I = f->info;
  if (I.dimensions == 0) return fail(TURDB_ERR_INVALID_ARGUMENT, "file header has dimensions == 0");
  const size_t dim = I.dimensions, n_total = I.n_nodes + I.n_tombstones;
  const bool need_copy = get_vector || present || I.n_tombstones;
  std::vector<float> full;
  const float* vsrc = vectors;
  if (need_copy) {
    try {
      full.resize(n_total * dim);
    } catch (...) {
      return fail(TURDB_ERR_OUT_OF_MEMORY, "cannot buffer %zu vectors", n_total);
    }
    const float inf = std::numeric_limits<float>::infinity();
    for (size_t i = 0; i < n_total; ++i) {
      float* dst = full.data() + i * dim;
      bool have = i < I.n_nodes;
      if (have && present) have = present[i] != 0;
      if (have && get_vector) have = get_vector(user, f->row_ids[i], dst) != 0;
      else if (have) memcpy(dst, vectors + i * dim, dim * 4);
      if (!have) std::fill(dst, dst + dim, inf);
    }
    vsrc = full.data();
  }
  turdb_cuda_graph g{};
  turdb_cuda_hnsw_file_graph(f, vsrc, &g);
  return turdb_cuda_index_create(&g, device, out);
}

}  // extern "C"
