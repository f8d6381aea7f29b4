// Trajectory files in the reference's on-disk format, written without libhdf5.
//
// run!(::SplittingMethod, h5file) stores the integrator state of every step in an HDF5 dataset
// "z" of Julia size (nd, np, nt+1), chunk (nd, np, 1), unlimited along time
// (src/methods/splitting.jl:32-34,41); run!(::GeometricIntegrator, h5file) stores "z" of size
// (np, nt+1), chunk (np, 1), and "t" of size (nt+1), chunk (1) (src/methods/geometric_integrator.jl:21-25,
// 34-35), and the scripts read them back with h5read (scripts/vlasov_poisson.jl:38,
// scripts/lenard_bernstein_conservative.jl:46).  HDF5.jl reverses Julia's column-major dimensions, so in
// the file "z" has the row-major shape (nt+1, np, nd) with chunks (1, np, nd): one chunk = one saved frame.
//
// This writer emits exactly that subset of the HDF5 file format (HDF5 File Format Specification 2.0):
// superblock version 0, a root group in the old symbol-table form (local heap, version-1 B-tree of type 0,
// one SNOD), version-1 object headers with dataspace (v1, with maximum dimensions), IEEE-754 little-endian
// double datatype (v1), fill-value (v2) and chunked data-layout (v3) messages, and per dataset a
// version-1 B-tree of type 1 indexing one unfiltered chunk per frame.  The number of frames is known when
// the run starts (the reference knows ntime() too), so every address is fixed at commit time: the metadata
// is written once, the file is sized, and frames are then written in place (pwrite), in any order and in
// pieces -- which is what lets the device-to-host copy of a frame stream straight into the file.  Unwritten
// frames read as zeros; the file is valid after every call.
//
// tests/test_h5_cpu.py reads these files back with an independent reader that is itself pinned on a file
// produced by the real HDF5 library.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "vpm_internal.h"

namespace {

constexpr uint64_t kUndef = ~0ull;
constexpr int kGroupLeafK = 4;       // superblock "group leaf node K": a SNOD holds up to 2K entries
constexpr int kGroupInternalK = 16;  // superblock "group internal node K"
constexpr int kChunkK = 32;          // indexed-storage internal node K (the library default; superblock v0 has no field)
constexpr int kMaxDatasets = 2 * kGroupLeafK;
constexpr int kMaxRank = 4;

struct Buf {
    std::vector<uint8_t> b;
    void u8(unsigned v) { b.push_back((uint8_t)v); }
    void le(uint64_t v, int n) { for (int i = 0; i < n; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void u16(uint64_t v) { le(v, 2); }
    void u32(uint64_t v) { le(v, 4); }
    void u64(uint64_t v) { le(v, 8); }
    void str(const char* s, size_t n) { b.insert(b.end(), s, s + n); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void pad_to(size_t n) { if (b.size() < n) zeros(n - b.size()); }
    void pad8() { while (b.size() % 8) b.push_back(0); }
    size_t size() const { return b.size(); }
};

uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct Dset {
    std::string name;
    int rank = 0;
    uint64_t dims[kMaxRank] = {};   // dims[0] = number of frames
    uint64_t chunk_bytes = 0;
    uint64_t heap_off = 0, ohdr = 0, btree = 0, data = 0;
    std::vector<std::vector<uint64_t>> levels;  // node addresses per B-tree level (0 = leaves)
};

}  // namespace

struct vpm_h5 {
    int fd = -1;
    std::string path;
    bool committed = false;
    std::vector<Dset> ds;
    uint64_t eof = 0;
    // opt-in (VPM_H5_THREADS > 1): the file is mapped and large writes are copied by several threads.  Buffered
    // pwrite()s to one file serialise on its inode lock, page faults on a shared mapping do not: 4x on tmpfs in a
    // micro-benchmark, nothing on a disk-backed file system, not yet measured on the GPU box -- hence off by default.
    // (A mapped sparse file turns "disk full" into SIGBUS instead of an error code; another reason for opt-in.)
    char* map = nullptr;
    int nthreads = 1;
};

namespace {

using vpm::fail;

size_t chunk_key_bytes(int rank) { return 8 + 8 * (size_t)(rank + 1); }
size_t chunk_node_bytes(int rank) { return 24 + 2 * kChunkK * 8 + (2 * kChunkK + 1) * chunk_key_bytes(rank); }

void put_chunk_key(Buf& o, const Dset& d, uint64_t frame, bool sentinel)
{
    o.u32(sentinel ? 0 : d.chunk_bytes);  // size of the stored chunk
    o.u32(0);                             // filter mask
    o.u64(frame);                         // offset along the frame axis (chunk extent 1)
    for (int i = 1; i <= d.rank; i++) o.u64(0);  // remaining axes + the element-size entry
}

// one B-tree node of type 1 (raw data chunks) covering children [first, first+count) of the level below
void put_chunk_node(Buf& o, const Dset& d, int level, size_t index, size_t first, size_t count, uint64_t span)
{
    const std::vector<uint64_t>& nodes = d.levels[level];
    const size_t start = o.size();
    o.str("TREE", 4);
    o.u8(1);
    o.u8(level);
    o.u16(count);
    o.u64(index > 0 ? nodes[index - 1] : kUndef);
    o.u64(index + 1 < nodes.size() ? nodes[index + 1] : kUndef);
    // span = number of frames below one child of this node; child c starts at frame (first + c) * span
    for (size_t c = 0; c < count; c++) {
        put_chunk_key(o, d, (first + c) * span, false);
        o.u64(level == 0 ? d.data + (first + c) * d.chunk_bytes : d.levels[level - 1][first + c]);
    }
    const uint64_t next = (first + count) * span;
    put_chunk_key(o, d, std::min<uint64_t>(next, d.dims[0]), next >= d.dims[0]);
    o.pad_to(start + chunk_node_bytes(d.rank));
}

int pwrite_all(int fd, const void* buf, size_t n, uint64_t off)
{
    const char* p = (const char*)buf;
    while (n) {
        ssize_t w = pwrite(fd, p, n, (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            return fail(VPM_ERR_INVALID, std::string("vpm_h5: write failed: ") + strerror(errno));
        }
        p += w; off += (uint64_t)w; n -= (size_t)w;
    }
    return VPM_OK;
}

}  // namespace

extern "C" {

int vpm_h5_create(const char* path, vpm_h5** out)
{
    if (!path || !out) return fail(VPM_ERR_INVALID, "vpm_h5_create: NULL argument");
    int fd = open(path, O_CREAT | O_TRUNC | O_RDWR, 0644);
    if (fd < 0) return fail(VPM_ERR_INVALID, std::string("vpm_h5_create: cannot open ") + path + ": " + strerror(errno));
    vpm_h5* f = new (std::nothrow) vpm_h5;
    if (!f) { close(fd); return fail(VPM_ERR_NOMEM, "vpm_h5_create: out of memory"); }
    f->fd = fd;
    f->path = path;
    *out = f;
    return VPM_OK;
}

int vpm_h5_add_dataset(vpm_h5* f, const char* name, int rank, const int64_t* dims, int* id_out)
{
    if (!f || !name || !dims) return fail(VPM_ERR_INVALID, "vpm_h5_add_dataset: NULL argument");
    if (f->committed) return fail(VPM_ERR_INVALID, "vpm_h5_add_dataset: the file layout is already committed");
    if (rank < 1 || rank > kMaxRank - 1) return fail(VPM_ERR_INVALID, "vpm_h5_add_dataset: rank must be 1..3");
    if ((int)f->ds.size() >= kMaxDatasets) return fail(VPM_ERR_UNSUPPORTED, "vpm_h5_add_dataset: at most 8 datasets per file");
    const size_t len = strlen(name);
    if (len == 0 || len > 255 || strchr(name, '/')) return fail(VPM_ERR_INVALID, "vpm_h5_add_dataset: bad dataset name");
    for (const Dset& d : f->ds)
        if (d.name == name) return fail(VPM_ERR_INVALID, "vpm_h5_add_dataset: duplicate dataset name");
    Dset d;
    d.name = name;
    d.rank = rank;
    unsigned __int128 bytes = 8;
    for (int i = 0; i < rank; i++) {
        if (dims[i] < 1) return fail(VPM_ERR_INVALID, "vpm_h5_add_dataset: dimensions must be positive");
        d.dims[i] = (uint64_t)dims[i];
        if (i > 0) bytes *= (uint64_t)dims[i];
    }
    // HDF5 stores a chunk's size in 32 bits; the reference's one-frame chunks have the same limit
    if (bytes >= ((unsigned __int128)1 << 32))
        return fail(VPM_ERR_UNSUPPORTED, "vpm_h5_add_dataset: one frame must be smaller than 4 GiB (HDF5 chunk size limit); "
                                         "write one file per particle slab");
    d.chunk_bytes = (uint64_t)bytes;
    f->ds.push_back(d);
    if (id_out) *id_out = (int)f->ds.size() - 1;
    return VPM_OK;
}

int vpm_h5_commit(vpm_h5* f)
{
    if (!f) return fail(VPM_ERR_INVALID, "vpm_h5_commit: NULL handle");
    if (f->committed) return VPM_OK;
    if (f->ds.empty()) return fail(VPM_ERR_INVALID, "vpm_h5_commit: no datasets");

    // ---- address plan ---------------------------------------------------------------------
    const uint64_t sb_size = 96, root_ohdr = sb_size, root_ohdr_size = 16 + 8 + 16;
    const uint64_t heap_hdr = root_ohdr + root_ohdr_size, heap_data = heap_hdr + 32;
    uint64_t names = 8;  // offset 0 holds the empty string (the root group's own name)
    for (Dset& d : f->ds) {
        d.heap_off = names;
        names += align_up(d.name.size() + 1, 8);
    }
    const uint64_t heap_size = align_up(names + 16, 8) + 64;  // names + one free block
    const uint64_t gtree = heap_data + heap_size;
    const uint64_t gtree_size = 24 + 2 * kGroupInternalK * 8 + (2 * kGroupInternalK + 1) * 8;
    const uint64_t snod = gtree + gtree_size, snod_size = 8 + 2 * kGroupLeafK * 40;
    uint64_t at = snod + snod_size;
    for (Dset& d : f->ds) {
        d.ohdr = at;
        const uint64_t space = 8 + 16 * (uint64_t)d.rank, layout = align_up(3 + 8 + 4 * (uint64_t)(d.rank + 1), 8);
        at += 16 + (8 + space) + (8 + 24) + (8 + 8) + (8 + layout);
    }
    for (Dset& d : f->ds) {
        d.levels.clear();
        size_t count = (size_t)d.dims[0];  // entries to index at this level
        do {
            const size_t nodes = (count + 2 * kChunkK - 1) / (2 * kChunkK);
            std::vector<uint64_t> addrs(nodes);
            for (size_t i = 0; i < nodes; i++) { addrs[i] = at; at += chunk_node_bytes(d.rank); }
            d.levels.push_back(addrs);
            count = nodes;
        } while (count > 1);
        d.btree = d.levels.back()[0];
    }
    // small datasets first, then the big frames on 4 KiB boundaries
    std::vector<Dset*> order;
    for (Dset& d : f->ds) order.push_back(&d);
    std::stable_sort(order.begin(), order.end(), [](const Dset* a, const Dset* b) { return a->chunk_bytes < b->chunk_bytes; });
    for (Dset* d : order) {
        at = align_up(at, d->chunk_bytes >= 4096 ? 4096 : 8);
        d->data = at;
        at += d->chunk_bytes * d->dims[0];
    }
    f->eof = at;

    // ---- metadata image -------------------------------------------------------------------
    Buf o;
    o.str("\x89HDF\r\n\x1a\n", 8);
    o.u8(0); o.u8(0); o.u8(0); o.u8(0);  // superblock, free-space, root symbol table versions; reserved
    o.u8(0); o.u8(8); o.u8(8); o.u8(0);  // shared header version, size of offsets, size of lengths, reserved
    o.u16(kGroupLeafK); o.u16(kGroupInternalK);
    o.u32(0);                            // file consistency flags
    o.u64(0); o.u64(kUndef); o.u64(f->eof); o.u64(kUndef);  // base, free-space info, end of file, driver info
    o.u64(0); o.u64(root_ohdr); o.u32(1); o.u32(0); o.u64(gtree); o.u64(heap_hdr);  // root symbol table entry
    // root object header: one symbol table message
    o.u8(1); o.u8(0); o.u16(1); o.u32(1); o.u32(8 + 16); o.u32(0);
    o.u16(0x11); o.u16(16); o.u8(0); o.zeros(3); o.u64(gtree); o.u64(heap_hdr);
    // local heap
    o.str("HEAP", 4); o.u8(0); o.zeros(3); o.u64(heap_size); o.u64(names); o.u64(heap_data);
    o.zeros(8);
    for (const Dset& d : f->ds) { o.str(d.name.c_str(), d.name.size() + 1); o.pad8(); }
    o.u64(1); o.u64(heap_size - names);  // the free block: next = 1 (none), size
    o.pad_to(gtree);
    // group B-tree: one leaf entry pointing at the SNOD
    std::vector<const Dset*> sorted;
    for (const Dset& d : f->ds) sorted.push_back(&d);
    std::sort(sorted.begin(), sorted.end(), [](const Dset* a, const Dset* b) { return a->name < b->name; });
    o.str("TREE", 4); o.u8(0); o.u8(0); o.u16(1); o.u64(kUndef); o.u64(kUndef);
    o.u64(0); o.u64(snod); o.u64(sorted.back()->heap_off);
    o.pad_to(snod);
    o.str("SNOD", 4); o.u8(1); o.u8(0); o.u16(sorted.size());
    for (const Dset* d : sorted) { o.u64(d->heap_off); o.u64(d->ohdr); o.u32(0); o.u32(0); o.zeros(16); }
    o.pad_to(snod + snod_size);
    for (const Dset& d : f->ds) {
        const uint64_t space = 8 + 16 * (uint64_t)d.rank, layout = align_up(3 + 8 + 4 * (uint64_t)(d.rank + 1), 8);
        o.pad_to(d.ohdr);
        o.u8(1); o.u8(0); o.u16(4); o.u32(1); o.u32((8 + space) + (8 + 24) + (8 + 8) + (8 + layout)); o.u32(0);
        // dataspace: version 1, maximum dimensions present, unlimited along the frame axis
        o.u16(0x01); o.u16(space); o.u8(0); o.zeros(3);
        o.u8(1); o.u8(d.rank); o.u8(1); o.zeros(5);
        for (int i = 0; i < d.rank; i++) o.u64(d.dims[i]);
        o.u64(kUndef);
        for (int i = 1; i < d.rank; i++) o.u64(d.dims[i]);
        // datatype: IEEE-754 binary64, little-endian
        o.u16(0x03); o.u16(24); o.u8(1); o.zeros(3);
        o.u8(0x11); o.u8(0x20); o.u8(63); o.u8(0); o.u32(8);
        o.u16(0); o.u16(64); o.u8(52); o.u8(11); o.u8(0); o.u8(52); o.u32(1023);
        o.zeros(4);
        // fill value: version 2, early allocation (every chunk exists), fill time "if set", default value (defined, size 0)
        // -- the combination libhdf5 itself writes for a dataset without a user fill value; unwritten frames read as 0
        o.u16(0x05); o.u16(8); o.u8(1); o.zeros(3);
        o.u8(2); o.u8(1); o.u8(2); o.u8(1); o.u32(0);
        // data layout: version 3, chunked, B-tree address, chunk dimensions + element size
        o.u16(0x08); o.u16(layout); o.u8(0); o.zeros(3);
        const size_t lstart = o.size();
        o.u8(3); o.u8(2); o.u8(d.rank + 1); o.u64(d.btree);
        o.u32(1);
        for (int i = 1; i < d.rank; i++) o.u32(d.dims[i]);
        o.u32(8);
        o.pad_to(lstart + layout);
    }
    for (const Dset& d : f->ds) {
        uint64_t span = 1;
        size_t below = (size_t)d.dims[0];
        for (size_t lv = 0; lv < d.levels.size(); lv++) {
            for (size_t i = 0; i < d.levels[lv].size(); i++) {
                o.pad_to(d.levels[lv][i]);
                const size_t first = i * 2 * kChunkK, count = std::min<size_t>(2 * kChunkK, below - first);
                put_chunk_node(o, d, (int)lv, i, first, count, span);
            }
            below = d.levels[lv].size();
            span *= 2 * kChunkK;
        }
    }
    if (ftruncate(f->fd, (off_t)f->eof) != 0)
        return fail(VPM_ERR_INVALID, std::string("vpm_h5_commit: cannot size ") + f->path + ": " + strerror(errno));
    int rc = pwrite_all(f->fd, o.b.data(), o.size(), 0);
    if (rc != VPM_OK) return rc;
    if (const char* e = std::getenv("VPM_H5_THREADS")) {
        const int hw = (int)std::thread::hardware_concurrency();
        f->nthreads = std::max(1, std::min(std::atoi(e), hw > 0 ? hw : 1));
        if (f->nthreads > 1) {
            void* m = mmap(nullptr, (size_t)f->eof, PROT_READ | PROT_WRITE, MAP_SHARED, f->fd, 0);
            if (m != MAP_FAILED) f->map = (char*)m;
            else f->nthreads = 1;   // fall back to pwrite
        }
    }
    f->committed = true;
    return VPM_OK;
}

int vpm_h5_write(vpm_h5* f, int id, int64_t frame, int64_t offset_doubles, int64_t count, const double* data)
{
    if (!f || !data) return fail(VPM_ERR_INVALID, "vpm_h5_write: NULL argument");
    if (!f->committed) return fail(VPM_ERR_INVALID, "vpm_h5_write: call vpm_h5_commit first");
    if (id < 0 || id >= (int)f->ds.size()) return fail(VPM_ERR_INVALID, "vpm_h5_write: bad dataset id");
    const Dset& d = f->ds[id];
    if (frame < 0 || (uint64_t)frame >= d.dims[0] || offset_doubles < 0 || count < 0 ||
        (uint64_t)(offset_doubles + count) * 8 > d.chunk_bytes)
        return fail(VPM_ERR_INVALID, "vpm_h5_write: frame or range outside the dataset");
    const uint64_t at = d.data + (uint64_t)frame * d.chunk_bytes + (uint64_t)offset_doubles * 8;
    const size_t bytes = (size_t)count * 8;
    if (!f->map) return pwrite_all(f->fd, data, bytes, at);
    constexpr size_t kSlice = (size_t)1 << 20;   // below 1 MiB per thread a plain copy wins
    const int nt = (int)std::min<size_t>((size_t)f->nthreads, bytes / kSlice);
    if (nt <= 1) {
        std::memcpy(f->map + at, data, bytes);
        return VPM_OK;
    }
    const size_t per = (bytes / (size_t)nt + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) {
        const size_t lo = std::min(bytes, (size_t)t * per), hi = std::min(bytes, lo + per);
        if (hi > lo) th.emplace_back([=] { std::memcpy(f->map + at + lo, (const char*)data + lo, hi - lo); });
    }
    for (std::thread& t : th) t.join();
    return VPM_OK;
}

int vpm_h5_close(vpm_h5* f)
{
    if (!f) return VPM_OK;
    int rc = VPM_OK;
    if (f->map) munmap(f->map, (size_t)f->eof);
    if (f->fd >= 0 && close(f->fd) != 0) rc = fail(VPM_ERR_INVALID, std::string("vpm_h5_close: ") + strerror(errno));
    delete f;
    return rc;
}

}  // extern "C"
