// Coordinate ingest in front of the contact-map kernel: host code only, no kernels.
//
// Reference: extract_calpha_coords (mDeepFRI/pdb.py:130-162) decompresses every FoldComp hit to PDB TEXT and hands it to
// extract_residues_coordinates (bio_utils.py:281-302) -> biotite PDBFile.read(...).get_structure()[0] -> chain "A", atom_name
// "CA", hetero False (bio_utils.py:230-255): milliseconds of Python per structure, three to four orders of magnitude slower than
// the GPU path consumes structures.  Two pieces replace it:
//   * mdf_pdb_calpha / mdf_pdb_calpha_batch: a column-slice parser of PDB text with biotite's selection rules (first model,
//     ATOM records only, one chain, atom name CA, first alternate location), multi-threaded over structures;
//   * the C-alpha cache: one mmap-able file per structure database written once (ids, float32 [L, 3] blocks, an id hash
//     table), whose lookups return POINTERS into the mapping - exactly the (pointer, rows) arrays mdf_path_submit_ragged takes.
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "mdf_common.cuh"

using namespace mdf;

namespace {

// three-letter residue names -> one-letter (biotite ProteinSequence: the 20 standard residues + B, Z, X, U, O, J)
char three_to_one(const char *r)
{
    static const char *tab[] = {"ALA", "A", "ARG", "R", "ASN", "N", "ASP", "D", "CYS", "C", "GLN", "Q", "GLU", "E", "GLY", "G", "HIS", "H",
                                "ILE", "I", "LEU", "L", "LYS", "K", "MET", "M", "PHE", "F", "PRO", "P", "SER", "S", "THR", "T", "TRP", "W",
                                "TYR", "Y", "VAL", "V", "ASX", "B", "GLX", "Z", "UNK", "X", "SEC", "U", "PYL", "O", "XLE", "J"};
    for (size_t i = 0; i < sizeof tab / sizeof *tab; i += 2)
        if (r[0] == tab[i][0] && r[1] == tab[i][1] && r[2] == tab[i][2]) return tab[i + 1][0];
    return 0;
}

// float(line[a:b]) the way Python parses it: surrounding blanks ignored, the value rounded to double, then stored as float32.
// Plain fixed-point fields ("%8.3f", what every PDB writer emits) take the fast path: the digits as one integer m (< 2^53) divided
// by 10^k - both exact in double, and IEEE division rounds correctly, so the result equals strtod's.  Anything else goes to strtod.
bool field_to_float(const char *p, int width, float *out)
{
    static const double p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    int a = 0, b = 0;
    while (b < width && p[b] != '\n' && p[b] != '\r' && p[b] != 0) ++b;
    while (a < b && p[a] == ' ') ++a;
    while (b > a && p[b - 1] == ' ') --b;
    if (a == b) return false;
    int i = a;
    bool neg = false;
    if (p[i] == '-' || p[i] == '+') { neg = p[i] == '-'; ++i; }
    unsigned long long m = 0;
    int digits = 0, frac = -1;
    for (; i < b; ++i) {
        const char c = p[i];
        if (c >= '0' && c <= '9') { m = m * 10 + (unsigned)(c - '0'); ++digits; if (frac >= 0) ++frac; }
        else if (c == '.' && frac < 0) frac = 0;
        else break;
    }
    if (i == b && digits > 0 && digits <= 15) {
        const double v = (double)m / p10[frac < 0 ? 0 : frac];
        *out = (float)(neg ? -v : v);
        return true;
    }
    char buf[24];
    const int n = std::min(b - a, 23);
    memcpy(buf, p + a, (size_t)n);
    buf[n] = 0;
    char *end = nullptr;
    const double v = strtod(buf, &end);
    if (end == buf || *end) return false;
    *out = (float)v;
    return true;
}

struct CaScan {
    int n = 0;              // selected C-alpha atoms
    int bad_line = -1;      // 1-based line of a malformed coordinate field
    bool chain_seen = false;
    char bad_res[4] = {0, 0, 0, 0};
};

// One pass over the text.  coords / res1 / res3 may be NULL (count only).  Selection = biotite's:
//   first model only (everything after the first ENDMDL is ignored); records starting with "ATOM" (HETATM = hetero);
//   chain_id == chain; atom_name.strip() == "CA"; alternate locations: the first one seen in each residue.
CaScan scan_pdb(const char *t, size_t len, char chain, float *coords, char *res1, char *res3, int capacity)
{
    CaScan r;
    const char *p = t, *end = t + len;
    int line_no = 0;
    // residue identity of the last atom, and the alternate location kept for it
    char cur_res[12] = {0};
    char kept_alt = ' ';
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        const size_t ll = (size_t)(le - p);
        ++line_no;
        if (ll >= 6 && !memcmp(p, "ENDMDL", 6)) break;
        const bool atom = ll >= 54 && !memcmp(p, "ATOM", 4), het = ll >= 54 && !memcmp(p, "HETATM", 6);
        if (atom || het) {
            if (p[21] == chain) r.chain_seen = true;
            // altloc bookkeeping follows every atom of the chain's residues, as biotite filters before the CA selection
            char res_id[12];
            memcpy(res_id, p + 17, 10);        // resName(17-19) chain(21) resSeq(22-25) iCode(26)
            res_id[10] = 0;
            if (memcmp(res_id, cur_res, 10) != 0) { memcpy(cur_res, res_id, 11); kept_alt = ' '; }
            const char alt = p[16];
            bool alt_ok = true;
            if (alt != ' ') {
                if (kept_alt == ' ') kept_alt = alt;
                alt_ok = alt == kept_alt;
            }
            if (atom && alt_ok && p[21] == chain) {
                // atom name = columns 13-16 stripped
                int a = 12, b = 16;
                while (a < b && p[a] == ' ') ++a;
                while (b > a && p[b - 1] == ' ') --b;
                if (b - a == 2 && p[a] == 'C' && p[a + 1] == 'A') {
                    if (coords || res1 || res3) {
                        if (r.n < capacity) {
                            if (coords) {
                                float *c = coords + (size_t)r.n * 3;
                                if (!field_to_float(p + 30, 8, c) || !field_to_float(p + 38, 8, c + 1) || !field_to_float(p + 46, 8, c + 2)) {
                                    if (r.bad_line < 0) r.bad_line = line_no;
                                }
                            }
                            if (res3) memcpy(res3 + (size_t)r.n * 3, p + 17, 3);
                            if (res1) {
                                const char one = three_to_one(p + 17);
                                res1[r.n] = one ? one : '?';
                                if (!one && !r.bad_res[0]) memcpy(r.bad_res, p + 17, 3);
                            }
                        }
                    }
                    ++r.n;
                }
            }
        }
        if (!nl) break;
        p = nl + 1;
    }
    return r;
}

void parallel_for(int n, int threads, const std::function<void(int, int)> &fn)
{
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    const int per = (n + threads - 1) / threads;
    for (int k = 0; k < threads; ++k) {
        const int lo = k * per, hi = std::min(n, lo + per);
        if (lo < hi) th.emplace_back(fn, lo, hi);
    }
    for (auto &t : th) t.join();
}

}  // namespace

// ---- bio_utils.py:281-302  extract_residues_coordinates(structure_string, chain, filetype="pdb") ------------------------------
extern "C" int mdf_pdb_calpha(const char *text, size_t len, char chain, float *coords, char *residues, char *resnames3, int capacity,
                              int *n_out)
{
    MDF_REQUIRE(text && n_out, "mdf_pdb_calpha: bad arguments");
    const CaScan r = scan_pdb(text, len, chain, coords, residues, resnames3, capacity);
    *n_out = r.n;
    MDF_REQUIRE(r.chain_seen, "Chain %c not found in structure.", chain);        // bio_utils.py:243-244
    if (!coords && !residues && !resnames3) return MDF_OK;
    MDF_REQUIRE(r.n <= capacity, "mdf_pdb_calpha: %d C-alpha atoms, capacity %d", r.n, capacity);
    MDF_REQUIRE(r.bad_line < 0, "mdf_pdb_calpha: malformed coordinate field on line %d", r.bad_line);
    MDF_REQUIRE(!residues || !r.bad_res[0], "non-standard residue %s", r.bad_res);
    return MDF_OK;
}

// n structures at once on `threads` host threads: rows[p] = C-alpha count of text p (-1: chain not found / malformed - such
// structures contribute no rows, like the reference's skipped alignments, pipeline.py:432-444); coords = flat float32 [sum rows, 3].
// Call with coords == NULL to size the buffer.
extern "C" int mdf_pdb_calpha_batch(int n, const char *const *texts, const int64_t *lens, char chain, int threads, int *rows, float *coords,
                                    int64_t capacity_rows, int64_t *total_rows)
{
    MDF_REQUIRE(n >= 0 && rows && total_rows && (n == 0 || (texts && lens)), "mdf_pdb_calpha_batch: bad arguments");
    parallel_for(n, threads, [&](int lo, int hi) {
        for (int p = lo; p < hi; ++p) {
            const CaScan r = scan_pdb(texts[p], (size_t)lens[p], chain, nullptr, nullptr, nullptr, 0);
            rows[p] = r.chain_seen ? r.n : -1;
        }
    });
    std::vector<int64_t> off((size_t)n + 1, 0);
    for (int p = 0; p < n; ++p) off[(size_t)p + 1] = off[(size_t)p] + std::max(rows[p], 0);
    *total_rows = off[(size_t)n];
    if (!coords) return MDF_OK;
    MDF_REQUIRE(capacity_rows >= *total_rows, "mdf_pdb_calpha_batch: capacity %lld rows < %lld", (long long)capacity_rows, (long long)*total_rows);
    std::vector<int> bad((size_t)n, 0);
    parallel_for(n, threads, [&](int lo, int hi) {
        for (int p = lo; p < hi; ++p) {
            if (rows[p] <= 0) continue;
            const CaScan r = scan_pdb(texts[p], (size_t)lens[p], chain, coords + off[(size_t)p] * 3, nullptr, nullptr, rows[p]);
            if (r.bad_line >= 0 || r.n != rows[p]) bad[(size_t)p] = 1;
        }
    });
    for (int p = 0; p < n; ++p) MDF_REQUIRE(!bad[(size_t)p], "mdf_pdb_calpha_batch: structure %d has a malformed coordinate field", p);
    return MDF_OK;
}

// ------------------------------------------------------------------------------------------- C-alpha cache
namespace {
const char kMagic[8] = {'M', 'D', 'F', 'C', 'A', '0', '0', '1'};
struct CacheHeader { char magic[8]; uint64_t n, total_rows, names_bytes, slots, row_off_at, name_off_at, names_at, hash_at, coords_at; };

uint64_t fnv1a(const char *s, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
    return h;
}
}  // namespace

struct mdf_coords_cache {
    int fd = -1;
    const char *map = nullptr;
    size_t bytes = 0;
    const CacheHeader *h = nullptr;
    const int64_t *row_off = nullptr, *name_off = nullptr;
    const char *names = nullptr;
    const uint32_t *hash = nullptr;
    const float *coords = nullptr;
};

extern "C" int mdf_coords_cache_create(const char *path, int64_t n, const char *const *ids, const int *rows, const float *const *coords)
{
    MDF_REQUIRE(path && n >= 0 && (n == 0 || (ids && rows && coords)), "mdf_coords_cache_create: bad arguments");
    CacheHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kMagic, 8);
    h.n = (uint64_t)n;
    std::vector<int64_t> row_off((size_t)n + 1, 0), name_off((size_t)n + 1, 0);
    for (int64_t p = 0; p < n; ++p) {
        MDF_REQUIRE(ids[p] && rows[p] >= 0 && (rows[p] == 0 || coords[p]), "mdf_coords_cache_create: entry %lld is incomplete", (long long)p);
        row_off[(size_t)p + 1] = row_off[(size_t)p] + rows[p];
        name_off[(size_t)p + 1] = name_off[(size_t)p] + (int64_t)strlen(ids[p]);
    }
    h.total_rows = (uint64_t)row_off[(size_t)n];
    h.names_bytes = (uint64_t)name_off[(size_t)n];
    uint64_t slots = 16;
    while (slots < (uint64_t)n * 2 + 1) slots <<= 1;
    h.slots = slots;
    std::vector<uint32_t> table((size_t)slots, 0u);
    for (int64_t p = 0; p < n; ++p) {
        const size_t len = (size_t)(name_off[(size_t)p + 1] - name_off[(size_t)p]);
        uint64_t s = fnv1a(ids[p], len) & (slots - 1);
        while (table[(size_t)s]) {
            const int64_t q = (int64_t)table[(size_t)s] - 1;
            const size_t ql = (size_t)(name_off[(size_t)q + 1] - name_off[(size_t)q]);
            MDF_REQUIRE(!(ql == len && !memcmp(ids[q], ids[p], len)), "mdf_coords_cache_create: duplicate id '%s'", ids[p]);
            s = (s + 1) & (slots - 1);
        }
        table[(size_t)s] = (uint32_t)(p + 1);
    }
    uint64_t at = sizeof(CacheHeader);
    auto place = [&](uint64_t bytes, uint64_t align) { at = (at + align - 1) / align * align; const uint64_t here = at; at += bytes; return here; };
    h.row_off_at = place((uint64_t)(n + 1) * 8, 64);
    h.name_off_at = place((uint64_t)(n + 1) * 8, 64);
    h.names_at = place(h.names_bytes, 64);
    h.hash_at = place(slots * 4, 64);
    h.coords_at = place(h.total_rows * 12, 4096);
    const std::string tmp = std::string(path) + ".tmp";
    FILE *fh = fopen(tmp.c_str(), "wb");
    if (!fh) { set_error("%s: cannot create: %s", tmp.c_str(), strerror(errno)); return MDF_ENOENT; }
    bool ok = true;
    auto put_at = [&](uint64_t where, const void *src, size_t bytes) {
        if (!bytes) return;
        ok = ok && fseek(fh, (long)where, SEEK_SET) == 0 && fwrite(src, 1, bytes, fh) == bytes;
    };
    put_at(0, &h, sizeof h);
    put_at(h.row_off_at, row_off.data(), row_off.size() * 8);
    put_at(h.name_off_at, name_off.data(), name_off.size() * 8);
    for (int64_t p = 0; p < n; ++p) put_at(h.names_at + (uint64_t)name_off[(size_t)p], ids[p], (size_t)(name_off[(size_t)p + 1] - name_off[(size_t)p]));
    put_at(h.hash_at, table.data(), table.size() * 4);
    for (int64_t p = 0; p < n; ++p) put_at(h.coords_at + (uint64_t)row_off[(size_t)p] * 12, coords[p], (size_t)rows[p] * 12);
    if (h.total_rows == 0) { const char z = 0; put_at(h.coords_at, &z, 1); }
    ok = fclose(fh) == 0 && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) { set_error("%s: write failed: %s", path, strerror(errno)); unlink(tmp.c_str()); return MDF_ENOENT; }
    return MDF_OK;
}

extern "C" int mdf_coords_cache_close(mdf_coords_cache *c)
{
    if (!c) return MDF_OK;
    if (c->map) munmap((void *)c->map, c->bytes);
    if (c->fd >= 0) close(c->fd);
    delete c;
    return MDF_OK;
}

extern "C" int mdf_coords_cache_open(const char *path, mdf_coords_cache **out)
{
    MDF_REQUIRE(path && out, "mdf_coords_cache_open: bad arguments");
    mdf_coords_cache *c = new mdf_coords_cache();
    c->fd = open(path, O_RDONLY);
    if (c->fd < 0) { set_error("%s: cannot open coordinate cache: %s", path, strerror(errno)); delete c; return MDF_ENOENT; }
    struct stat st;
    if (fstat(c->fd, &st) != 0 || (size_t)st.st_size < sizeof(CacheHeader)) { set_error("%s: not a C-alpha cache (too short)", path); mdf_coords_cache_close(c); return MDF_EPARSE; }
    c->bytes = (size_t)st.st_size;
    void *m = mmap(nullptr, c->bytes, PROT_READ, MAP_SHARED, c->fd, 0);
    if (m == MAP_FAILED) { c->map = nullptr; set_error("%s: mmap failed: %s", path, strerror(errno)); mdf_coords_cache_close(c); return MDF_ENOENT; }
    c->map = (const char *)m;
    c->h = (const CacheHeader *)c->map;
    const CacheHeader &h = *c->h;
    const bool sane = !memcmp(h.magic, kMagic, 8) && h.slots && !(h.slots & (h.slots - 1)) && h.row_off_at + (h.n + 1) * 8 <= c->bytes &&
                      h.name_off_at + (h.n + 1) * 8 <= c->bytes && h.names_at + h.names_bytes <= c->bytes && h.hash_at + h.slots * 4 <= c->bytes &&
                      h.coords_at + h.total_rows * 12 <= c->bytes;
    if (!sane) { set_error("%s: not a C-alpha cache written by mdf_coords_cache_create", path); mdf_coords_cache_close(c); return MDF_EPARSE; }
    c->row_off = (const int64_t *)(c->map + h.row_off_at);
    c->name_off = (const int64_t *)(c->map + h.name_off_at);
    c->names = c->map + h.names_at;
    c->hash = (const uint32_t *)(c->map + h.hash_at);
    c->coords = (const float *)(c->map + h.coords_at);
    if ((uint64_t)c->row_off[h.n] != h.total_rows || (uint64_t)c->name_off[h.n] != h.names_bytes) {
        set_error("%s: corrupt C-alpha cache (offset tables)", path);
        mdf_coords_cache_close(c);
        return MDF_EPARSE;
    }
    madvise((void *)c->map, c->bytes, MADV_WILLNEED);
    *out = c;
    return MDF_OK;
}

extern "C" int64_t mdf_coords_cache_size(const mdf_coords_cache *c) { return c ? (int64_t)c->h->n : 0; }

// coords_out[p] -> float32 [rows_out[p], 3] inside the mapping (valid until close), or NULL / -1 for an unknown id.
// Returns the number of ids NOT found through *missing (may be NULL).
extern "C" int mdf_coords_cache_lookup(const mdf_coords_cache *c, int64_t n, const char *const *ids, const float **coords_out, int *rows_out,
                                       int64_t *missing)
{
    MDF_REQUIRE(c && n >= 0 && (n == 0 || (ids && coords_out && rows_out)), "mdf_coords_cache_lookup: bad arguments");
    const uint64_t mask = c->h->slots - 1;
    int64_t miss = 0;
    for (int64_t p = 0; p < n; ++p) {
        coords_out[p] = nullptr;
        rows_out[p] = -1;
        if (!ids[p]) { ++miss; continue; }
        const size_t len = strlen(ids[p]);
        uint64_t s = fnv1a(ids[p], len) & mask;
        for (;;) {
            const uint32_t e = c->hash[s];
            if (!e) { ++miss; break; }
            const int64_t q = (int64_t)e - 1;
            if ((size_t)(c->name_off[q + 1] - c->name_off[q]) == len && !memcmp(c->names + c->name_off[q], ids[p], len)) {
                coords_out[p] = c->coords + c->row_off[q] * 3;
                rows_out[p] = (int)(c->row_off[q + 1] - c->row_off[q]);
                break;
            }
            s = (s + 1) & mask;
        }
    }
    if (missing) *missing = miss;
    return MDF_OK;
}

// id of entry q (NUL-terminated copy into buf) and its row count: enumeration for tools / tests
extern "C" int mdf_coords_cache_entry(const mdf_coords_cache *c, int64_t q, char *buf, size_t capacity, int *rows)
{
    MDF_REQUIRE(c && q >= 0 && (uint64_t)q < c->h->n && buf, "mdf_coords_cache_entry: bad arguments");
    const size_t len = (size_t)(c->name_off[q + 1] - c->name_off[q]);
    MDF_REQUIRE(len + 1 <= capacity, "mdf_coords_cache_entry: buffer too small");
    memcpy(buf, c->names + c->name_off[q], len);
    buf[len] = 0;
    if (rows) *rows = (int)(c->row_off[q + 1] - c->row_off[q]);
    return MDF_OK;
}
