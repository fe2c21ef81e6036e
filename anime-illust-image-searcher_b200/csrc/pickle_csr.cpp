// Streaming reader for the reference's `bm25_corpus` file (genmodel.py:84-85: `pickle.dump(bm25_corpus, f)`, a Python
// list of N dicts {term id: tf}) straight into doc-major CSR arrays - without materialising N Python dicts, which is what
// load_model() pays on every cold start (webui.py:680: tens of GB of Python objects at 10^7 docs).
//
// The file is an ordinary pickle (protocol 2..5).  A list of dicts of small ints uses a tiny opcode subset:
//   PROTO FRAME EMPTY_LIST EMPTY_DICT MEMOIZE/BINPUT/LONG_BINPUT MARK BININT1 BININT2 BININT LONG1 SETITEM SETITEMS
//   APPEND APPENDS STOP
// and its grammar is flat: every EMPTY_DICT opens the next document, every integer after it alternates key, value.  Any
// other opcode (numpy scalars as keys, nested containers ...) makes the reader give up with AIS_ERR_UNSUPPORTED so the
// caller can fall back to `pickle.load`.  Two passes: ais_pickle_csr_scan counts docs / entries, ais_pickle_csr_fill
// writes row_ptr [n+1], term_ids [nnz], tfs [nnz] (dict insertion order = the order genmodel.py:64-66 met the tags).
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/ais_b200.h"

namespace {

struct Mapped {
    const unsigned char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open_file(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); fd = -1; return false; }
        n = (size_t)st.st_size;
        if (n == 0) { p = nullptr; return true; }
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { ::close(fd); fd = -1; return false; }
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const unsigned char*)m;
        return true;
    }
    ~Mapped() {
        if (p) munmap((void*)p, n);
        if (fd >= 0) ::close(fd);
    }
};

// walks the opcode stream; Sink gets doc boundaries and integers
template <typename Sink>
int walk(const unsigned char* p, size_t n, Sink& sink, char* err, size_t err_len) {
    size_t i = 0;
    bool seen_list = false, stopped = false;
    auto need = [&](size_t k) { return i + k <= n; };
    while (i < n && !stopped) {
        const unsigned char op = p[i++];
        switch (op) {
            case 0x80: if (!need(1)) goto truncated; i += 1; break;                      // PROTO
            case 0x95: if (!need(8)) goto truncated; i += 8; break;                      // FRAME
            case ']': if (!seen_list) seen_list = true; else goto unsupported; break;    // EMPTY_LIST (only the outer one)
            case '}': if (!seen_list) goto unsupported; sink.open_doc(); break;          // EMPTY_DICT
            case 0x94: break;                                                            // MEMOIZE
            case 'q': if (!need(1)) goto truncated; i += 1; break;                       // BINPUT
            case 'r': if (!need(4)) goto truncated; i += 4; break;                       // LONG_BINPUT
            case '(': case 'u': case 's': case 'e': case 'a': break;                     // MARK SETITEMS SETITEM APPENDS APPEND
            case 'K': if (!need(1)) goto truncated; if (!sink.integer((int64_t)p[i])) goto unsupported; i += 1; break;
            case 'M': if (!need(2)) goto truncated; if (!sink.integer((int64_t)(p[i] | (p[i + 1] << 8)))) goto unsupported; i += 2; break;
            case 'J': {
                if (!need(4)) goto truncated;
                int32_t v;
                memcpy(&v, p + i, 4);
                if (!sink.integer((int64_t)v)) goto unsupported;
                i += 4;
                break;
            }
            case 0x8a: {                                                                 // LONG1: little-endian two's complement
                if (!need(1)) goto truncated;
                const unsigned len = p[i++];
                if (len > 8 || !need(len)) goto unsupported;
                int64_t v = 0;
                for (unsigned k = 0; k < len; ++k) v |= (int64_t)p[i + k] << (8 * k);
                if (len > 0 && len < 8 && (p[i + len - 1] & 0x80)) v -= (int64_t)1 << (8 * len);
                if (!sink.integer(v)) goto unsupported;
                i += len;
                break;
            }
            case '.': stopped = true; break;                                             // STOP
            default: goto unsupported;
        }
        continue;
    unsupported:
        snprintf(err, err_len, "opcode 0x%02x at byte %zu is outside the list-of-int-dicts subset", (unsigned)p[i - 1], i - 1);
        return AIS_ERR_UNSUPPORTED;
    truncated:
        snprintf(err, err_len, "pickle truncated at byte %zu", i);
        return AIS_ERR_INVALID;
    }
    if (!stopped || !seen_list) {
        snprintf(err, err_len, "not a pickled list (no STOP / no outer list)");
        return AIS_ERR_INVALID;
    }
    return sink.finish() ? AIS_OK : AIS_ERR_INVALID;
}

struct CountSink {
    int64_t docs = 0, ints = 0;
    bool in_doc = false;
    void open_doc() { ++docs; in_doc = true; }
    bool integer(int64_t) { if (!in_doc) return false; ++ints; return true; }
    bool finish() { return (ints & 1) == 0; }
};

struct FillSink {
    int64_t* row_ptr; int32_t* term_ids; int32_t* tfs;
    int64_t docs = 0, at = 0;
    bool in_doc = false, have_key = false;
    int64_t key = 0;
    bool range_ok = true;
    void open_doc() { row_ptr[docs++] = at; in_doc = true; }
    bool integer(int64_t v) {
        if (!in_doc) return false;
        if (!have_key) { key = v; have_key = true; return true; }
        if (key < INT32_MIN || key > INT32_MAX || v < INT32_MIN || v > INT32_MAX) range_ok = false;
        term_ids[at] = (int32_t)key;
        tfs[at] = (int32_t)v;
        ++at;
        have_key = false;
        return true;
    }
    bool finish() { row_ptr[docs] = at; return !have_key && range_ok; }
};

thread_local char g_perr[256] = "";

}  // namespace

extern "C" {

const char* ais_pickle_last_error(void) { return g_perr; }

int ais_pickle_csr_scan(const char* path, int64_t* out_n_docs, int64_t* out_nnz) {
    if (!path || !out_n_docs || !out_nnz) { snprintf(g_perr, sizeof(g_perr), "NULL argument"); return AIS_ERR_INVALID; }
    Mapped m;
    if (!m.open_file(path)) { snprintf(g_perr, sizeof(g_perr), "cannot open %s", path); return AIS_ERR_INVALID; }
    CountSink s;
    const int st = walk(m.p, m.n, s, g_perr, sizeof(g_perr));
    if (st != AIS_OK) return st;
    *out_n_docs = s.docs;
    *out_nnz = s.ints / 2;
    return AIS_OK;
}

int ais_pickle_csr_fill(const char* path, int64_t n_docs, int64_t nnz, int64_t* row_ptr, int32_t* term_ids, int32_t* tfs) {
    if (!path || !row_ptr || (nnz > 0 && (!term_ids || !tfs))) { snprintf(g_perr, sizeof(g_perr), "NULL argument"); return AIS_ERR_INVALID; }
    Mapped m;
    if (!m.open_file(path)) { snprintf(g_perr, sizeof(g_perr), "cannot open %s", path); return AIS_ERR_INVALID; }
    // the counts of the scan pass bound every write below
    CountSink c;
    int st = walk(m.p, m.n, c, g_perr, sizeof(g_perr));
    if (st != AIS_OK) return st;
    if (c.docs != n_docs || c.ints / 2 != nnz) {
        snprintf(g_perr, sizeof(g_perr), "file holds %lld docs / %lld entries, caller sized for %lld / %lld", (long long)c.docs,
                 (long long)(c.ints / 2), (long long)n_docs, (long long)nnz);
        return AIS_ERR_INVALID;
    }
    FillSink f{row_ptr, term_ids, tfs};
    st = walk(m.p, m.n, f, g_perr, sizeof(g_perr));
    if (st == AIS_ERR_INVALID && !f.range_ok) snprintf(g_perr, sizeof(g_perr), "a term id or tf does not fit 32 bits");
    return st;
}

}  // extern "C"
