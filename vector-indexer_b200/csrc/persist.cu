// persist.cu -- cold-start path: the reference's on-disk formats.
//   shards/shard_{id}.bin   Shard::save_to / get_centroid_vectors_from  (src/shards.rs:22-51, :68-177, :188-349)
//   index/index.bin         IvfIndex::save_to / load_index_from         (src/ivf_index.rs:274-316)
// Files written here are meant to be readable by the reference and vice versa.  The
// shard layout is fully specified by the reference's #[repr(C)] structs.  index.bin is
// bincode 2.0.1 `standard()` (little endian, varint integers) over serde; the ndarray
// 0.15.6 serde wrapper {v: u8 = 1, dim, data} is restated from the published crate
// (sources are not on disk) and is flagged "unverified" in DESIGN.md.
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <sys/stat.h>

#include "../../include/vidx_b200.h"
#include "index.h"

namespace vidx {

namespace {

#pragma pack(push, 1)
struct ShardHeader {  // shards.rs:22-31 (40 bytes; the source comment says 48)
    uint64_t shard_id, version;
    uint32_t dimensions, num_centroids;
    uint64_t index_offset, data_offset;
};
struct CentroidIndexEntry {  // shards.rs:34-42
    uint64_t centroid_id;
    uint32_t num_vectors, padding;
    uint64_t data_offset, data_size;
};
struct VectorMetaRec {  // shards.rs:45-51
    uint64_t id, external_id, timestamp;
};
#pragma pack(pop)
static_assert(sizeof(ShardHeader) == 40 && sizeof(CentroidIndexEntry) == 32 && sizeof(VectorMetaRec) == 24, "layout");

void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty()) mkdir(cur.c_str(), 0755);
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
}

// ---- bincode 2 varint ------------------------------------------------------------------
void put_varint(std::vector<uint8_t>& o, uint64_t v) {
    if (v < 251) {
        o.push_back((uint8_t)v);
    } else if (v < (1ull << 16)) {
        o.push_back(251);
        for (int i = 0; i < 2; i++) o.push_back((uint8_t)(v >> (8 * i)));
    } else if (v < (1ull << 32)) {
        o.push_back(252);
        for (int i = 0; i < 4; i++) o.push_back((uint8_t)(v >> (8 * i)));
    } else {
        o.push_back(253);
        for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i)));
    }
}
struct Reader {
    const uint8_t* p;
    size_t n, pos = 0;
    void need(size_t k) {
        if (pos + k > n) throw ApiError(VIDX_ERR_OTHER, "Bincode decoding error: unexpected end of index.bin");
    }
    uint8_t u8() { need(1); return p[pos++]; }
    uint64_t le(int bytes) {
        need(bytes);
        uint64_t v = 0;
        for (int i = 0; i < bytes; i++) v |= (uint64_t)p[pos + i] << (8 * i);
        pos += bytes;
        return v;
    }
    uint64_t varint() {
        uint8_t b = u8();
        if (b < 251) return b;
        if (b == 251) return le(2);
        if (b == 252) return le(4);
        if (b == 253) return le(8);
        throw ApiError(VIDX_ERR_OTHER, "Bincode decoding error: unsupported integer width");
    }
    float f32() {
        uint32_t v = (uint32_t)le(4);
        float f;
        std::memcpy(&f, &v, 4);
        return f;
    }
};

}  // namespace

// ---- save --------------------------------------------------------------------------------
namespace {
// fwrite that cannot fail silently (a full disk must not leave a truncated shard behind a success code)
struct CheckedFile {
    FILE* f = nullptr;
    std::string path;
    CheckedFile(const std::string& p) : path(p) {
        remove(p.c_str());
        f = fopen(p.c_str(), "wb");
        if (!f) throw ApiError(VIDX_ERR_OTHER, "cannot create " + p + ": " + strerror(errno));
    }
    void write(const void* src, size_t bytes) {
        if (bytes && fwrite(src, 1, bytes, f) != bytes) {
            const std::string why = strerror(errno);
            fclose(f);
            f = nullptr;
            throw ApiError(VIDX_ERR_OTHER, "short write to " + path + ": " + why);
        }
    }
    void close() {
        FILE* g = f;
        f = nullptr;
        if (g && fclose(g) != 0) throw ApiError(VIDX_ERR_OTHER, "write failed: " + path);
    }
    ~CheckedFile() { if (f) fclose(f); }
};
}  // namespace

// A rank of a shard-partitioned index writes the shard files it owns; index.bin comes from rank 0 (every rank
// holds the same centroid table).
void save_index(const Index& ix, const std::vector<float>& host_vectors /* by build position */,
                const std::string& index_dir, const std::string& shards_dir) {
    const uint32_t D = ix.dim;
    mkdirs(shards_dir);
    const size_t vsz = (size_t)D * 4, pad = (8 - vsz % 8) % 8;
    // shard files: lists in ascending list id (ivf_index.rs:152-164)
    std::vector<std::vector<uint32_t>> shard_lists(ix.num_shards);
    for (uint64_t l = 0; l < ix.nlist; l++) {
        if (ix.c2shard[l] >= shard_lists.size()) shard_lists.resize(ix.c2shard[l] + 1);
        shard_lists[ix.c2shard[l]].push_back((uint32_t)l);
    }
    const char zeros[8] = {0};
    for (size_t s = 0; s < shard_lists.size(); s++) {
        const auto& ls = shard_lists[s];
        if (ix.resident_partial) {  // not this rank's shard
            bool mine = true;
            for (uint32_t l : ls) mine = mine && ix.list_fully_resident(l);
            if (!mine) continue;
        }
        CheckedFile f(shards_dir + "/shard_" + std::to_string(s) + ".bin");
        ShardHeader hd{(uint64_t)s, 1, D, (uint32_t)ls.size(), 40, 40 + 32ull * ls.size()};
        f.write(&hd, sizeof hd);
        uint64_t off = hd.data_offset;
        for (uint32_t l : ls) {
            uint64_t size = vsz + pad + (uint64_t)ix.list_len[l] * (24 + vsz + pad);
            CentroidIndexEntry e{l, ix.list_len[l], 0, off, size};
            f.write(&e, sizeof e);
            off += size;
        }
        for (uint32_t l : ls) {
            f.write(&ix.centroids[(size_t)l * D], vsz);
            f.write(zeros, pad);
            uint64_t r0 = (uint64_t)ix.res_g0[l] * kGroup;
            for (uint32_t j = 0; j < ix.list_len[l]; j++) {
                uint32_t src = ix.row_src[r0 + j];
                VectorMetaRec m{ix.internal_ids.empty() ? (uint64_t)src : ix.internal_ids[src], ix.ext_ids[src], ix.timestamps[src]};
                f.write(&m, sizeof m);
                f.write(&host_vectors[(size_t)src * D], vsz);
                f.write(zeros, pad);
            }
        }
        f.close();
    }
    if (ix.resident_partial && ix.part_rank != 0) return;
    write_index_bin(index_dir, ix.centroids.data(), ix.c2shard.data(), ix.nlist, D);
}

// index.bin = bincode 2.0.1 `standard()` (little endian, varint integers) over serde of
//   IvfIndex { centroids: Array1<Centroid{id: usize, vector: Vec<f32>}>, centroids_to_shard: Array1<usize>, dimension: u32 }
// (ivf_index.rs:36-41, :274-294); ndarray 0.15.6 serialises an array as {v: u8 = 1, dim, data: seq}.
void write_index_bin(const std::string& index_dir, const float* centroids, const uint32_t* c2shard, uint64_t nlist, uint32_t D) {
    const size_t vsz = (size_t)D * 4;
    std::vector<uint8_t> o;
    o.push_back(1);  // ndarray ARRAY_FORMAT_VERSION
    put_varint(o, nlist);  // dim: [usize; 1]
    put_varint(o, nlist);  // data: seq length
    for (uint64_t l = 0; l < nlist; l++) {
        put_varint(o, l);  // Centroid.id
        put_varint(o, D);  // Vec<f32> length
        const uint8_t* b = reinterpret_cast<const uint8_t*>(centroids + (size_t)l * D);
        o.insert(o.end(), b, b + vsz);
    }
    o.push_back(1);
    put_varint(o, nlist);
    put_varint(o, nlist);
    for (uint64_t l = 0; l < nlist; l++) put_varint(o, c2shard[l]);
    put_varint(o, D);  // dimension: u32
    mkdirs(index_dir);
    CheckedFile f(index_dir + "/index.bin");
    f.write(o.data(), o.size());
    f.close();
}

// ---- load --------------------------------------------------------------------------------
static uint64_t file_size(FILE* f) {
    struct stat st;
    return fstat(fileno(f), &st) == 0 ? (uint64_t)st.st_size : 0;
}
// Phase 1: index.bin, then header + centroid index of every shard file.  Every size read from a file is checked
// against the file's length before anything is allocated from it.  A shard that cannot be opened or does not parse
// is skipped as a whole -- search_with_paths drops failed shard reads (ivf_index.rs:254) -- and reported in
// skipped_shards; its lists stay empty.
void read_index_bin(const std::string& index_dir, LoadedMeta& out) {
    std::string ipath = index_dir + "/index.bin";
    FILE* f = fopen(ipath.c_str(), "rb");
    if (!f) throw ApiError(errno == ENOENT ? VIDX_ERR_NOT_FOUND : VIDX_ERR_OTHER, "cannot open " + ipath + ": " + strerror(errno));
    std::vector<uint8_t> buf;
    {
        uint8_t tmp[1 << 16];
        size_t g;
        while ((g = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + g);
        fclose(f);
    }
    Reader r{buf.data(), buf.size()};
    if (r.u8() != 1) throw ApiError(VIDX_ERR_OTHER, "Bincode decoding error: unknown array version");
    uint64_t n1 = r.varint(), n2 = r.varint();
    if (n1 != n2) throw ApiError(VIDX_ERR_OTHER, "Bincode decoding error: array shape mismatch");
    if (n1 > buf.size()) throw ApiError(VIDX_ERR_OTHER, "Bincode decoding error: array longer than the file");
    out.nlist = n1;
    out.centroids.clear();
    uint64_t D = 0;
    for (uint64_t l = 0; l < n1; l++) {
        (void)r.varint();  // Centroid.id
        uint64_t len = r.varint();
        if (l == 0) D = len;
        if (len != D) throw ApiError(VIDX_ERR_OTHER, "Bincode decoding error: ragged centroid vectors");
        r.need(len * 4);
        for (uint64_t d = 0; d < len; d++) out.centroids.push_back(r.f32());
    }
    if (r.u8() != 1) throw ApiError(VIDX_ERR_OTHER, "Bincode decoding error: unknown array version");
    uint64_t m1 = r.varint(), m2 = r.varint();
    if (m1 != m2 || m1 != n1) throw ApiError(VIDX_ERR_OTHER, "Bincode decoding error: centroids_to_shard shape mismatch");
    out.c2shard.resize(n1);
    uint64_t max_shard = 0;
    for (uint64_t l = 0; l < n1; l++) {
        out.c2shard[l] = (uint32_t)r.varint();
        max_shard = std::max<uint64_t>(max_shard, out.c2shard[l]);
    }
    out.dim = (uint32_t)r.varint();
    if (n1 && D != out.dim) throw ApiError(VIDX_ERR_INVALID_DATA, "index.bin: centroid length differs from dimension");
    if (max_shard > n1) throw ApiError(VIDX_ERR_INVALID_DATA, "index.bin: shard id larger than the number of lists");
    out.num_shards = n1 ? max_shard + 1 : 0;
}

void load_index_meta(const std::string& index_dir, const std::string& shards_dir, LoadedMeta& out) {
    read_index_bin(index_dir, out);
    const uint64_t n1 = out.nlist;
    const size_t vsz = (size_t)out.dim * 4, pad = (8 - vsz % 8) % 8;
    out.list_len.assign(n1, 0);
    out.list_file_shard.assign(n1, 0);
    out.list_block_off.assign(n1, 0);
    out.skipped_shards.clear();
    for (uint64_t s = 0; s < out.num_shards; s++) {
        std::string path = shards_dir + "/shard_" + std::to_string(s) + ".bin";
        FILE* sf = fopen(path.c_str(), "rb");
        auto skip = [&](const char* why) {
            out.skipped_shards.push_back(path + ": " + why);
            if (sf) fclose(sf);
        };
        if (!sf) { skip("cannot open"); continue; }
        const uint64_t fsz = file_size(sf);
        ShardHeader hd;
        if (fread(&hd, sizeof hd, 1, sf) != 1) { skip("invalid shard header"); continue; }
        if (hd.shard_id != s) { skip("shard id mismatch"); continue; }
        if (hd.dimensions != out.dim) { skip("dimension mismatch"); continue; }
        if (hd.index_offset > fsz || (uint64_t)hd.num_centroids * sizeof(CentroidIndexEntry) > fsz - hd.index_offset) {
            skip("invalid index");
            continue;
        }
        std::vector<CentroidIndexEntry> ents(hd.num_centroids);
        if (fseek(sf, (long)hd.index_offset, SEEK_SET) != 0 ||
            (hd.num_centroids && fread(ents.data(), sizeof(CentroidIndexEntry), hd.num_centroids, sf) != hd.num_centroids)) {
            skip("invalid index");
            continue;
        }
        // the whole shard is staged and committed only when every entry is sane
        bool bad = false;
        for (const auto& e : ents) {
            const uint64_t want = vsz + pad + (uint64_t)e.num_vectors * (24 + vsz + pad);
            if (e.centroid_id >= n1 || e.data_offset > fsz || e.data_size > fsz - e.data_offset || e.data_size < want) {
                bad = true;
                break;
            }
        }
        if (bad) { skip("invalid cluster block"); continue; }
        for (const auto& e : ents) {
            out.list_len[e.centroid_id] = e.num_vectors;
            out.list_file_shard[e.centroid_id] = (uint32_t)s;
            out.list_block_off[e.centroid_id] = e.data_offset;
        }
        fclose(sf);
    }
}

// Phase 2: vectors [v0[l], v1[l]) of every list (records are fixed-size, so a range is one seek + one read).
void load_list_ranges(const std::string& shards_dir, const LoadedMeta& m, const std::vector<uint32_t>& v0,
                      const std::vector<uint32_t>& v1, std::vector<float>& data, std::vector<uint64_t>& meta) {
    const size_t vsz = (size_t)m.dim * 4, pad = (8 - vsz % 8) % 8, rec = 24 + vsz + pad;
    uint64_t total = 0;
    for (uint64_t l = 0; l < m.nlist; l++) total += v1[l] - v0[l];
    data.resize((size_t)total * m.dim);
    meta.resize((size_t)total * 3);
    uint64_t at = 0;
    FILE* sf = nullptr;
    int64_t open_shard = -1;
    std::vector<uint8_t> blk;
    // lists grouped by file so that every shard file is opened once
    std::vector<uint32_t> order(m.nlist);
    for (uint64_t l = 0; l < m.nlist; l++) order[l] = (uint32_t)l;
    // (output order is list order; the file of consecutive lists changes rarely, reopen on change)
    for (uint32_t l : order) {
        const uint32_t n = v1[l] - v0[l];
        if (!n) continue;
        if (open_shard != (int64_t)m.list_file_shard[l]) {
            if (sf) fclose(sf);
            const std::string path = shards_dir + "/shard_" + std::to_string(m.list_file_shard[l]) + ".bin";
            sf = fopen(path.c_str(), "rb");
            if (!sf) throw ApiError(VIDX_ERR_OTHER, "cannot open " + path + ": " + strerror(errno));
            open_shard = m.list_file_shard[l];
        }
        blk.resize((size_t)n * rec);
        const uint64_t off = m.list_block_off[l] + vsz + pad + (uint64_t)v0[l] * rec;
        if (fseek(sf, (long)off, SEEK_SET) != 0 || fread(blk.data(), 1, blk.size(), sf) != blk.size()) {
            fclose(sf);
            throw ApiError(VIDX_ERR_OTHER, "short read in shard_" + std::to_string(open_shard) + ".bin");
        }
        for (uint32_t j = 0; j < n; j++, at++) {
            VectorMetaRec mr;
            std::memcpy(&mr, blk.data() + (size_t)j * rec, 24);
            meta[3 * at] = mr.id;
            meta[3 * at + 1] = mr.external_id;
            meta[3 * at + 2] = mr.timestamp;
            std::memcpy(&data[(size_t)at * m.dim], blk.data() + (size_t)j * rec + 24, vsz);
        }
    }
    if (sf) fclose(sf);
}

// ---- vector files (src/utils.rs:34-107) ----------------------------------------------------
// A vector file is a concatenation of bincode-2 `standard()` encodings of Vec<(u64, Vec<f32>, u64)> (batches of 1000 in
// generate_test_vectors_parallel, utils.rs:34-79): varint length, then per record varint id, varint vector length, f32 LE
// values, varint metadata.  read_vectors_from_file (utils.rs:82-107) decodes batch after batch and stops -- silently -- at the
// first batch that fails to decode.
void read_vector_file(const std::string& path, std::vector<uint64_t>& ids, std::vector<uint64_t>& lens, std::vector<float>& values,
                      std::vector<uint64_t>& meta) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw ApiError(VIDX_ERR_OTHER, "read_vectors_from_file: " + path + ": " + strerror(errno));
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(f);
    Reader r{buf.data(), buf.size()};
    while (r.pos < r.n) {
        const size_t n_ids = ids.size(), n_vals = values.size();
        try {
            const uint64_t cnt = r.varint();
            for (uint64_t i = 0; i < cnt; i++) {
                const uint64_t id = r.varint();
                const uint64_t len = r.varint();
                r.need(len * 4);
                for (uint64_t j = 0; j < len; j++) values.push_back(r.f32());
                const uint64_t m = r.varint();
                ids.push_back(id);
                lens.push_back(len);
                meta.push_back(m);
            }
        } catch (const ApiError&) {  // Err(_) => break: drop the partial batch
            ids.resize(n_ids);
            lens.resize(n_ids);
            meta.resize(n_ids);
            values.resize(n_vals);
            break;
        }
    }
}
void write_vector_file(const std::string& path, const float* data, const uint64_t* ids, const uint64_t* meta, uint64_t n, uint64_t dim,
                       uint64_t batch) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw ApiError(VIDX_ERR_OTHER, "cannot create " + path + ": " + strerror(errno));
    if (!batch) batch = 1000;
    for (uint64_t b0 = 0; b0 < n; b0 += batch) {
        const uint64_t b1 = std::min(n, b0 + batch);
        std::vector<uint8_t> o;
        put_varint(o, b1 - b0);
        for (uint64_t i = b0; i < b1; i++) {
            put_varint(o, ids ? ids[i] : i);
            put_varint(o, dim);
            const uint8_t* pv = reinterpret_cast<const uint8_t*>(data + i * dim);
            o.insert(o.end(), pv, pv + dim * 4);
            put_varint(o, meta ? meta[i] : 0);
        }
        if (fwrite(o.data(), 1, o.size(), f) != o.size()) {
            fclose(f);
            throw ApiError(VIDX_ERR_OTHER, "short write to " + path);
        }
    }
    if (fclose(f) != 0) throw ApiError(VIDX_ERR_OTHER, "write failed: " + path);
}

}  // namespace vidx
