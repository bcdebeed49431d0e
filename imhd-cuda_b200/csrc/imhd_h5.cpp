// Minimal HDF5 serializer for the reference's output contract -- no libhdf5 (none in
// the image).  Writes what HDF5 1.14 writes with default property lists for these calls:
//   fluidvars_<it>.h5  src/on-device/utils/phdf5_write_all.cpp:87-169   8 one-dimensional fp32 datasets
//                      rho, rhovx, rhovy, rhovz, Bx, By, Bz, e of Nx*Ny*Nz elements in IDX3D order
//   (frame 0 only)     src/on-device/utils/hdf5_write_attributes.cpp:55-98  cubeDimensions int32[3],
//                      cubeDimensionsNames vlen-string[3], storagePattern vlen-string[1] on every dataset
//   grid.h5            src/on-device/utils/hdf5_write_grid.cpp:78-137   x_grid, y_grid, z_grid + scalar
//                      attributes spacing (fp32) and dimension (int32)
// File structure ("version 0" family of the HDF5 File Format Specification): superblock v0, root group =
// v1 object header with a Symbol Table message -> v1 B-tree leaf + local heap + one symbol-table node whose
// entries are sorted by name, one v1 object header per dataset (Dataspace v1, Datatype v1, Fill Value v2,
// Layout v3 contiguous, Attribute v1 messages), one global heap collection for the vlen strings, raw data
// 8-byte aligned after the metadata.  tests/test_h5.py parses the result with an independent reader.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/imhd_b200.h"

namespace imhd { void set_error(const char* fmt, ...); }

namespace {

constexpr uint64_t UNDEF = ~0ull;

struct Buf {
    std::vector<uint8_t> b;
    size_t size() const { return b.size(); }
    void u8(uint8_t v) { b.push_back(v); }
    void u16(uint16_t v) { for (int i = 0; i < 2; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    void u32(uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void pad8() { while (b.size() % 8) b.push_back(0); }
    void append(const Buf& o) { b.insert(b.end(), o.b.begin(), o.b.end()); }
    void set_u64(size_t at, uint64_t v) { for (int i = 0; i < 8; ++i) b[at + i] = (uint8_t)(v >> (8 * i)); }
};

// ---- datatype messages (padded to 8) ----------------------------------------------------------------------
Buf dt_f32() {
    Buf m;
    m.u8(0x11); m.u8(0x20); m.u8(0x1F); m.u8(0x00);  // v1 class 1 (float); LE, implied-1 mantissa; sign bit 31
    m.u32(4);
    m.u16(0); m.u16(32); m.u8(23); m.u8(8); m.u8(0); m.u8(23); m.u32(127);
    m.pad8();
    return m;
}
Buf dt_i32() {
    Buf m;
    m.u8(0x10); m.u8(0x08); m.u8(0); m.u8(0);  // v1 class 0 (fixed point); LE, signed
    m.u32(4);
    m.u16(0); m.u16(32);
    m.pad8();
    return m;
}
Buf dt_vlen_str() {
    Buf m;
    m.u8(0x19); m.u8(0x01); m.u8(0x00); m.u8(0x00);  // v1 class 9 (vlen); type = string, null-terminated, ASCII
    m.u32(16);                                        // 4 (length) + 8 (heap address) + 4 (index)
    m.u8(0x10); m.u8(0x00); m.u8(0); m.u8(0);        // base type: 1-byte unsigned fixed point
    m.u32(1);
    m.u16(0); m.u16(8);
    m.pad8();
    return m;
}
Buf ds_simple(int rank, uint64_t dim) {  // Dataspace v1
    Buf m;
    m.u8(1); m.u8((uint8_t)rank); m.u8(0); m.u8(0); m.u32(0);
    if (rank == 1) m.u64(dim);
    return m;
}

struct Msg { uint16_t type; Buf data; };

Msg attribute(const char* name, const Buf& dt, const Buf& ds, const void* data, size_t nbytes) {
    Msg a;
    a.type = 0x000C;
    const size_t nlen = strlen(name) + 1;
    a.data.u8(1); a.data.u8(0);
    a.data.u16((uint16_t)nlen); a.data.u16((uint16_t)dt.size()); a.data.u16((uint16_t)ds.size());
    a.data.bytes(name, nlen); a.data.pad8();
    a.data.append(dt); a.data.pad8();
    a.data.append(ds); a.data.pad8();
    a.data.bytes(data, nbytes);
    a.data.pad8();
    return a;
}

// ---- global heap collection for vlen strings ------------------------------------------------------------------
struct GlobalHeap {
    std::vector<std::string> objs;
    uint32_t add(const std::string& s) { objs.push_back(s); return (uint32_t)objs.size(); }  // 1-based index
    Buf build() const {
        Buf g;
        g.bytes("GCOL", 4); g.u8(1); g.zeros(3);
        const size_t size_at = g.size();
        g.u64(0);
        for (size_t n = 0; n < objs.size(); ++n) {
            g.u16((uint16_t)(n + 1)); g.u16(1); g.u32(0);
            g.u64(objs[n].size() + 1);  // with the terminating NUL, as H5T_VARIABLE C strings are stored
            g.bytes(objs[n].c_str(), objs[n].size() + 1);
            g.pad8();
        }
        size_t total = g.size() + 16;
        if (total < 4096) total = 4096;  // H5HG_MINSIZE
        g.u16(0); g.u16(0); g.u32(0);
        g.u64(total - (g.size() - 8));   // free-space object: size includes its own 16-byte header
        g.zeros(total - g.size());
        g.set_u64(size_at, total);
        return g;
    }
};

struct Dataset {
    std::string name;
    const float* data;
    uint64_t n;
    std::vector<Msg> attrs;
};

Buf object_header(const std::vector<Msg>& msgs) {
    Buf body;
    for (const Msg& m : msgs) {
        body.u16(m.type); body.u16((uint16_t)m.data.size()); body.u8(0); body.zeros(3);
        body.append(m.data);
    }
    Buf h;
    h.u8(1); h.u8(0); h.u16((uint16_t)msgs.size()); h.u32(1); h.u32((uint32_t)body.size()); h.u32(0);
    h.append(body);
    return h;
}

// Assemble and write a file with the given datasets.  vlen attribute payloads reference the global heap by
// address, so the heap address is fixed first (metadata sizes do not depend on it).
int emit(const char* path, std::vector<Dataset>& dsets, const GlobalHeap& gh,
         const std::vector<std::pair<int, std::vector<std::pair<std::string, std::vector<uint32_t>>>>>& vl_attrs) {
    // names must be sorted: libhdf5 binary-searches the symbol table node
    for (size_t a = 0; a + 1 < dsets.size(); ++a)
        if (!(dsets[a].name < dsets[a + 1].name)) { imhd::set_error("h5: dataset names not sorted"); return IMHD_E_INVALID; }
    if (dsets.size() > 8) { imhd::set_error("h5: more than 8 datasets need a second symbol-table node"); return IMHD_E_INVALID; }

    // local heap data segment: "" at offset 0, then the names
    Buf heapdata;
    heapdata.zeros(8);
    std::vector<uint64_t> name_off;
    for (auto& d : dsets) { name_off.push_back(heapdata.size()); heapdata.bytes(d.name.c_str(), d.name.size() + 1); heapdata.pad8(); }

    // fixed-size metadata blocks, laid out back to back
    const uint64_t a_super = 0, a_root = 96, a_btree = a_root + 40, a_heap = a_btree + 544;
    const uint64_t a_heapdata = a_heap + 32, a_snod = a_heapdata + heapdata.size(), a_hdr0 = a_snod + 328;

    // two passes: header sizes first (with dummy addresses), then final addresses
    uint64_t a_gcol = 0;
    std::vector<uint64_t> a_hdr(dsets.size()), a_data(dsets.size());
    std::vector<Buf> hdrs(dsets.size());
    for (int pass = 0; pass < 2; ++pass) {
        uint64_t at = a_hdr0;
        for (size_t n = 0; n < dsets.size(); ++n) {
            std::vector<Msg> msgs;
            msgs.push_back({0x0001, ds_simple(1, dsets[n].n)});
            msgs.push_back({0x0003, dt_f32()});
            Buf fill; fill.u8(2); fill.u8(2); fill.u8(2); fill.u8(1); fill.u32(0);  // v2: late alloc, write if set, default fill (size 0)
            msgs.push_back({0x0005, fill});
            Buf lay; lay.u8(3); lay.u8(1); lay.u64(a_data[n]); lay.u64(dsets[n].n * 4); lay.pad8();
            msgs.push_back({0x0008, lay});
            for (const Msg& a : dsets[n].attrs) msgs.push_back(a);
            for (const auto& va : vl_attrs)
                if (va.first == (int)n || va.first < 0)
                    for (const auto& one : va.second) {
                        Buf payload;
                        for (uint32_t idx : one.second) {
                            payload.u32((uint32_t)gh.objs[idx - 1].size() + 1);
                            payload.u64(a_gcol);
                            payload.u32(idx);
                        }
                        msgs.push_back(attribute(one.first.c_str(), dt_vlen_str(), ds_simple(1, one.second.size()),
                                                 payload.b.data(), payload.size()));
                    }
            hdrs[n] = object_header(msgs);
            a_hdr[n] = at;
            at += hdrs[n].size();
        }
        a_gcol = at;
        const uint64_t gsize = gh.objs.empty() ? 0 : gh.build().size();
        uint64_t dat = a_gcol + gsize;
        for (size_t n = 0; n < dsets.size(); ++n) { a_data[n] = dat; dat += (dsets[n].n * 4 + 7) / 8 * 8; }
    }
    uint64_t eof = a_data.empty() ? a_gcol : a_data.back() + (dsets.back().n * 4 + 7) / 8 * 8;

    Buf f;
    // superblock v0
    const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    f.bytes(sig, 8);
    f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(8); f.u8(8); f.u8(0);
    f.u16(4); f.u16(16); f.u32(0);
    f.u64(0); f.u64(UNDEF); f.u64(eof); f.u64(UNDEF);
    f.u64(0); f.u64(a_root); f.u32(1); f.u32(0); f.u64(a_btree); f.u64(a_heap);  // root symbol-table entry (cached)
    // root group object header: Symbol Table message
    {
        Buf st; st.u64(a_btree); st.u64(a_heap);
        f.append(object_header({{0x0011, st}}));
    }
    // B-tree v1 leaf (group node): 1 entry
    f.bytes("TREE", 4); f.u8(0); f.u8(0); f.u16(1); f.u64(UNDEF); f.u64(UNDEF);
    f.u64(0); f.u64(a_snod); f.u64(name_off.empty() ? 0 : name_off.back());
    f.zeros(544 - (f.size() - a_btree));
    // local heap
    f.bytes("HEAP", 4); f.u8(0); f.zeros(3); f.u64(heapdata.size()); f.u64(1 /* H5HL_FREE_NULL */); f.u64(a_heapdata);
    f.append(heapdata);
    // symbol table node
    f.bytes("SNOD", 4); f.u8(1); f.u8(0); f.u16((uint16_t)dsets.size());
    for (size_t n = 0; n < 8; ++n) {
        if (n < dsets.size()) { f.u64(name_off[n]); f.u64(a_hdr[n]); f.u32(0); f.u32(0); f.zeros(16); }
        else f.zeros(40);
    }
    for (auto& h : hdrs) f.append(h);
    if (!gh.objs.empty()) f.append(gh.build());
    if (f.size() != (a_data.empty() ? a_gcol : a_data[0])) { imhd::set_error("h5: internal layout error"); return IMHD_E_IO; }
    (void)a_super;

    FILE* fp = fopen(path, "wb");
    if (!fp) { imhd::set_error("h5: cannot open %s for writing", path); return IMHD_E_IO; }
    bool ok = fwrite(f.b.data(), 1, f.size(), fp) == f.size();
    static const uint8_t zeros8[8] = {0};
    for (size_t n = 0; ok && n < dsets.size(); ++n) {
        const size_t nb = dsets[n].n * 4;
        ok = fwrite(dsets[n].data, 1, nb, fp) == nb;
        if (ok && nb % 8) ok = fwrite(zeros8, 1, 8 - nb % 8, fp) == 8 - nb % 8;
    }
    ok = (fclose(fp) == 0) && ok;
    if (!ok) { imhd::set_error("h5: short write to %s", path); return IMHD_E_IO; }
    return 0;
}

}  // namespace

extern "C" int imhd_h5_write_fluidvars(const char* path, const float* host_Q, int Nx, int Ny, int Nz, int with_attributes) {
    if (!path || !host_Q || Nx < 1 || Ny < 1 || Nz < 1) { imhd::set_error("imhd_h5_write_fluidvars: bad argument"); return IMHD_E_INVALID; }
    const uint64_t cube = (uint64_t)Nx * Ny * Nz;
    // dataset v is host_Q + v*cube (phdf5_write_all.cpp:136-169); symbol-table order is by name
    static const struct { const char* name; int v; } order[8] = {{"Bx", 4}, {"By", 5}, {"Bz", 6}, {"e", 7}, {"rho", 0}, {"rhovx", 1}, {"rhovy", 2}, {"rhovz", 3}};
    std::vector<Dataset> ds;
    for (auto& o : order) ds.push_back({o.name, host_Q + (uint64_t)o.v * cube, cube, {}});
    GlobalHeap gh;
    std::vector<std::pair<int, std::vector<std::pair<std::string, std::vector<uint32_t>>>>> vl;
    if (with_attributes) {
        // hdf5_write_attributes.cpp:62,71-72 hands an hsize_t[3] to H5Awrite as H5T_NATIVE_INT, so the three stored
        // int32 are the low/high words {Nx, 0, Ny} on a little-endian host.  Reproduced.
        const int32_t cd[3] = {Nx, 0, Ny};
        for (auto& d : ds) d.attrs.push_back(attribute("cubeDimensions", dt_i32(), ds_simple(1, 3), cd, sizeof(cd)));
        const uint32_t iNx = gh.add("Nx"), iNy = gh.add("Ny"), iNz = gh.add("Nz");
        const uint32_t iSP = gh.add("Row-major, depth-minor: l = k * (Nx * Ny) + i * Ny + j");
        vl.push_back({-1, {{"cubeDimensionsNames", {iNx, iNy, iNz}}, {"storagePattern", {iSP}}}});
    }
    return emit(path, ds, gh, vl);
}

extern "C" int imhd_h5_write_grid(const char* path, const float* x, const float* y, const float* z, int Nx, int Ny, int Nz) {
    if (!path || !x || !y || !z || Nx < 2 || Ny < 2 || Nz < 2) { imhd::set_error("imhd_h5_write_grid: bad argument"); return IMHD_E_INVALID; }
    std::vector<Dataset> ds = {{"x_grid", x, (uint64_t)Nx, {}}, {"y_grid", y, (uint64_t)Ny, {}}, {"z_grid", z, (uint64_t)Nz, {}}};
    const int32_t n[3] = {Nx, Ny, Nz};
    const float* g[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
        const float spacing = (g[a][n[a] - 1] - g[a][0]) / (n[a] - 1);  // hdf5_write_grid.cpp:91-93
        ds[a].attrs.push_back(attribute("spacing", dt_f32(), ds_simple(0, 0), &spacing, 4));
        ds[a].attrs.push_back(attribute("dimension", dt_i32(), ds_simple(0, 0), &n[a], 4));
    }
    GlobalHeap gh;
    return emit(path, ds, gh, {});
}
