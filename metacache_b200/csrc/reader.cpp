// FASTA / FASTQ (+gz) reader feeding query batches.  Host code only.
//
// Reference behaviour restated (nothing copied):
//   sequence_reader::read_next        sequence_io.cpp:157-226   record grammar, recovery at the next
//                                                               '>' / '@' line, multi-line data,
//                                                               one quality line, '\r' stripped
//   sequence_pair_reader              sequence_io.cpp:253-330   none / files (lockstep) / sequences
//   char_istream (gzopen for all)     sequence_iostream.hpp:221-234, 411-434
//   query_batched reader thread       database_query.hpp:257-281 one thread fills batches
// The reference pulls characters through a 64 KB buffer with one reader thread.  Here a reader
// parses records in place in a multi-megabyte buffer (memchr per line, zero copies for single-line
// sequences) and appends them straight to the pinned host buffers of a batch slot; any number of
// readers can work on disjoint byte ranges of one uncompressed file, one per host thread / slot.
#include "../../include/mcb200.h"

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>
#include <zlib.h>

extern "C" int mcb200_internal_set_error (int code, const char* msg);     // api.cu: fills mcb200_last_error()
#define mcb200_set_error mcb200_internal_set_error

namespace {

struct Record {
    const char* hdr = nullptr; size_t hlen = 0;      // header without the leading '>' / '@'
    const char* seq = nullptr; size_t len = 0;       // sequence characters, line breaks removed
};

struct Input {
    int fd = -1;
    gzFile gz = nullptr;
    std::vector<char> buf;
    size_t beg = 0, end = 0;
    bool eof = false;
    uint64_t buf_pos = 0;          // file offset of buf[0] (plain files)
    uint64_t limit = ~0ull;        // records starting at or after this file offset are not ours
    std::string scratch;           // multi-line sequences are joined here
    std::string io_error;          // set by refill() when the source failed (eof stays false)

    bool open (const char* path, std::string& err) {
        const int f = ::open(path, O_RDONLY);
        if (f < 0) { err = std::string("can't open file ") + path; return false; }
        unsigned char magic[2] = {0, 0};
        const ssize_t n = ::pread(f, magic, 2, 0);
        if (n == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            ::close(f);
            gz = gzopen(path, "rb");
            if (!gz) { err = std::string("can't open file ") + path; return false; }
            gzbuffer(gz, 1u << 20);
        } else fd = f;
        buf.resize(size_t(8) << 20);
        return true;
    }
    void close () {
        if (fd >= 0) ::close(fd);
        if (gz) gzclose(gz);
        fd = -1; gz = nullptr;
    }
    bool seek (uint64_t pos) {      // plain files only
        if (fd < 0) return false;
        if (::lseek(fd, off_t(pos), SEEK_SET) < 0) return false;
        beg = end = 0; buf_pos = pos; eof = false;
        return true;
    }
    // keeps [beg, end), moves it to the front, reads more; false when nothing could be added
    bool refill () {
        if (eof) return false;
        if (beg > 0) {
            memmove(buf.data(), buf.data() + beg, end - beg);
            buf_pos += beg; end -= beg; beg = 0;
        }
        if (end == buf.size()) buf.resize(buf.size() * 2);       // one record larger than the buffer
        const size_t want = buf.size() - end;
        long got;
        if (gz) got = gzread(gz, buf.data() + end, unsigned(std::min<size_t>(want, 1u << 30)));
        else    got = long(::read(fd, buf.data() + end, want));
        if (got < 0) {                                            // I/O error: NOT an end of input
            io_error = gz ? "error while decompressing the input" : (std::string("read failed: ") + strerror(errno));
            return false;
        }
        if (got == 0) {
            if (gz) {                                             // a truncated / corrupt stream ends with an error state
                int zerr = Z_OK;
                const char* msg = gzerror(gz, &zerr);
                if (zerr != Z_OK && zerr != Z_STREAM_END) {
                    io_error = std::string("corrupt or truncated gzip input: ") + (msg ? msg : "");
                    return false;
                }
            }
            eof = true;
            return false;
        }
        end += size_t(got);
        return true;
    }
    uint64_t offset_of (const char* p) const { return buf_pos + uint64_t(p - buf.data()); }
};

inline const char* find_nl (const char* p, const char* e) {
    return static_cast<const char*>(memchr(p, '\n', size_t(e - p)));
}

// One record from contiguous memory [p, e).  Returns bytes consumed (> 0), 0 if the record is not
// complete yet (and !eof), or -1 if there is no further record.
long parse_record (const char* p, const char* e, bool eof, Record& r, std::string& scratch, uint64_t* start_off)
{
    const char* const p0 = p;
    // recovery: skip lines until one starts with '>' or '@'  (sequence_io.cpp:165-170)
    while (p < e && *p != '>' && *p != '@') {
        const char* nl = find_nl(p, e);
        if (!nl) { if (!eof) return 0; p = e; break; }
        p = nl + 1;
    }
    if (p >= e) return eof ? -1 : 0;
    if (start_off) *start_off = uint64_t(p - p0);
    // header line
    const char* hs = p + 1;
    const char* he = find_nl(hs, e);
    if (!he) { if (!eof) return 0; he = e; }
    r.hdr = hs; r.hlen = size_t(he - hs);
    if (r.hlen && r.hdr[r.hlen - 1] == '\r') --r.hlen;
    r.seq = nullptr; r.len = 0;
    const char* q = (he < e) ? he + 1 : e;
    bool joined = false;
    for (;;) {
        if (q >= e) { if (!eof) return 0; return long(e - p0); }
        const char c = *q;
        if (c == '>') return long(q - p0);                         // next FASTA record
        if (c == '+') break;                                       // FASTQ separator line
        if (c == '\n') { ++q; continue; }                          // empty line
        const char* le = find_nl(q, e);
        const char* next;
        if (!le) { if (!eof) return 0; le = e; next = e; } else next = le + 1;
        size_t n = size_t(le - q);
        if (n && q[n - 1] == '\r') --n;
        if (r.len == 0 && !joined) { r.seq = q; r.len = n; }
        else {
            if (!joined) { scratch.assign(r.seq, r.len); joined = true; }
            scratch.append(q, n);
            r.seq = scratch.data(); r.len = scratch.size();
        }
        q = next;
    }
    // FASTQ: rest of the '+' line, then ONE quality line  (sequence_io.cpp:203-223)
    const char* nl = find_nl(q, e);
    if (!nl) { if (!eof) return 0; return long(e - p0); }
    q = nl + 1;
    if (q >= e) { if (!eof) return 0; return long(e - p0); }
    nl = find_nl(q, e);
    if (!nl) { if (!eof) return 0; return long(e - p0); }
    return long(nl + 1 - p0);
}

struct Source {
    Input in;
    uint64_t index = 0;
    bool next (Record& r, std::string& err) {
        for (;;) {
            uint64_t skipped = 0;
            const long used = (in.end > in.beg)
                ? parse_record(in.buf.data() + in.beg, in.buf.data() + in.end, in.eof, r, in.scratch, &skipped)
                : (in.eof ? -1 : 0);
            if (used > 0) {
                if (in.offset_of(in.buf.data() + in.beg) + skipped >= in.limit) return false;   // next range's record
                in.beg += size_t(used);
                ++index;
                return true;
            }
            if (used < 0) return false;
            if (!in.refill() && !in.eof) { err = in.io_error.empty() ? std::string("read error") : in.io_error; return false; }
        }
    }
};

} // namespace

struct mcb200_reader {
    Source a, b;
    int mode = 0;                      // 0 none, 1 files, 2 sequences
    bool pending = false;              // a parsed pair waits for room in a batch slot
    Record r1, r2;
    std::string hold1, hold2, holdh;   // copies that survive the next parse / refill
    uint64_t pairs = 0;
};

static mcb200_reader* reader_fail (mcb200_reader* r, int code, const std::string& msg) {
    mcb200_set_error(code, msg.c_str());
    if (r) { r->a.in.close(); r->b.in.close(); delete r; }
    return nullptr;
}

static bool check_first_char (Source& s, std::string& err) {
    if (s.in.end == s.in.beg) s.in.refill();
    if (s.in.end == s.in.beg || (s.in.buf[s.in.beg] != '>' && s.in.buf[s.in.beg] != '@')) {
        err = "malformed fasta/fastq file - expected header char '>' or '@' not found";
        return false;
    }
    return true;
}

extern "C" mcb200_reader* mcb200_reader_open (const char* filename1, const char* filename2) {
    if (!filename1 || !*filename1) return reader_fail(nullptr, MCB200_EINVAL, "no filename was given");
    mcb200_reader* r = new mcb200_reader;
    std::string err;
    if (!r->a.in.open(filename1, err)) return reader_fail(r, MCB200_EIO, err);
    if (!check_first_char(r->a, err)) return reader_fail(r, MCB200_EIO, err);
    if (filename2 && *filename2) {
        if (strcmp(filename1, filename2) != 0) {
            r->mode = 1;
            if (!r->b.in.open(filename2, err)) return reader_fail(r, MCB200_EIO, err);
            if (!check_first_char(r->b, err)) return reader_fail(r, MCB200_EIO, err);
        } else r->mode = 2;
    }
    return r;
}

// First record that starts at or after file offset `pos` of an uncompressed file.  FASTA: a line
// starting with '>'.  FASTQ: a line starting with '@' is a header or a quality line (sequence and
// '+' lines never start with '@'); the grammar has exactly ONE quality line per record
// (sequence_io.cpp:203-223) and it is followed by the next header, while a header is followed by a
// sequence or '+' line.  So "'@' line followed by another '@' line" = quality line, the record
// starts at the second one; any other '@' line is a header.  Multi-line sequences are fine.
static bool sync_to_record (Input& in, uint64_t pos, bool fastq, std::string& err) {
    if (pos == 0) return in.seek(0);
    if (!in.seek(pos - 1)) { err = "seek failed"; return false; }
    size_t scan = 0;                       // buffer offset the search for the next line break starts at
    auto more = [&] () -> bool {           // false: nothing further (end of file, or an I/O error left in `err`)
        if (in.refill()) return true;
        if (!in.eof) err = in.io_error.empty() ? std::string("read error") : in.io_error;
        return false;
    };
    for (;;) {                             // (data stays at buf[0]: beg == 0 throughout)
        if (in.end == scan && !more()) { in.beg = in.end; return err.empty(); }
        const char* base = in.buf.data();
        const char* e = base + in.end;
        const char* nl = find_nl(base + scan, e);
        if (!nl) {
            scan = in.end;
            if (in.eof) { in.beg = in.end; return true; }
            continue;
        }
        const size_t l0 = size_t(nl + 1 - base);
        if (l0 >= in.end) {                // the line break is the last byte we have
            if (!more()) { in.beg = in.end; return err.empty(); }
            continue;                      // same line break again, now with data behind it
        }
        if (!fastq) {
            if (base[l0] == '>') { in.beg = l0; return true; }
        } else if (base[l0] == '@') {
            const char* n0 = find_nl(base + l0, e);
            if (!n0 || n0 + 1 >= e) {      // need the first character of the following line
                if (more()) continue;
                if (!err.empty()) return false;
                // '@' line that is the last line of the file: in a well-formed file that is the quality
                // line of the record before (a header is always followed by at least one more line)
                in.beg = in.end;
                return true;
            }
            in.beg = (n0[1] == '@') ? size_t(n0 + 1 - base) : l0;
            return true;
        }
        scan = l0;
    }
}

extern "C" mcb200_reader* mcb200_reader_open_range (const char* filename, uint64_t byte_begin, uint64_t byte_end) {
    if (!filename || !*filename) return reader_fail(nullptr, MCB200_EINVAL, "no filename was given");
    if (byte_end < byte_begin) return reader_fail(nullptr, MCB200_EINVAL, "empty byte range");
    mcb200_reader* r = new mcb200_reader;
    std::string err;
    if (!r->a.in.open(filename, err)) return reader_fail(r, MCB200_EIO, err);
    if (r->a.in.gz) return reader_fail(r, MCB200_EINVAL, "byte ranges need an uncompressed file");
    if (!check_first_char(r->a, err)) return reader_fail(r, MCB200_EIO, err);
    const bool fastq = r->a.in.buf[r->a.in.beg] == '@';
    if (!sync_to_record(r->a.in, byte_begin, fastq, err)) return reader_fail(r, MCB200_EIO, err);
    r->a.in.limit = byte_end;
    return r;
}

extern "C" void mcb200_reader_close (mcb200_reader* r) {
    if (!r) return;
    r->a.in.close(); r->b.in.close();
    delete r;
}

extern "C" uint64_t mcb200_reader_index (const mcb200_reader* r) { return r ? r->pairs : 0; }

// next query (read or read pair) into r->r1 / r->r2; false at the end of the input
static bool next_pair (mcb200_reader* r, std::string& err) {
    if (!r->a.next(r->r1, err)) return false;
    r->r2 = Record{};
    if (r->mode == 1) {
        if (!r->b.next(r->r2, err)) return false;                  // lockstep: stops with the shorter file
    } else if (r->mode == 2) {
        // the second record is parsed from the same buffer: the first must survive a refill
        r->hold1.assign(r->r1.seq ? r->r1.seq : "", r->r1.len);
        r->holdh.assign(r->r1.hdr ? r->r1.hdr : "", r->r1.hlen);
        Record second;
        const bool have = r->a.next(second, err);
        r->r1.seq = r->hold1.data(); r->r1.hdr = r->holdh.data();
        if (have) r->r2 = second;
    }
    ++r->pairs;
    return true;
}

extern "C" int mcb200_reader_next (mcb200_reader* r, const char** header, uint64_t* header_len,
                                   const char** seq1, uint64_t* len1, const char** seq2, uint64_t* len2) {
    if (!r) return mcb200_set_error(MCB200_EINVAL, "null reader");
    std::string err;
    if (!r->pending && !next_pair(r, err)) {
        if (!err.empty()) return mcb200_set_error(MCB200_EIO, err.c_str());
        return 0;
    }
    r->pending = false;
    if (header) *header = r->r1.hdr;
    if (header_len) *header_len = r->r1.hlen;
    if (seq1) *seq1 = r->r1.seq;
    if (len1) *len1 = r->r1.len;
    if (seq2) *seq2 = r->r2.seq;
    if (len2) *len2 = r->r2.len;
    return 1;
}

extern "C" int64_t mcb200_reader_skip (mcb200_reader* r, uint64_t n, uint64_t* bases) {
    if (!r) return mcb200_set_error(MCB200_EINVAL, "null reader");
    std::string err;
    uint64_t done = 0, nb = 0;
    while (done < n) {
        if (!r->pending && !next_pair(r, err)) {
            if (!err.empty()) return mcb200_set_error(MCB200_EIO, err.c_str());
            break;
        }
        r->pending = false;
        nb += r->r1.len + r->r2.len;
        ++done;
    }
    if (bases) *bases = nb;
    return int64_t(done);
}

extern "C" int64_t mcb200_reader_fill_batch (mcb200_reader* r, mcb200_batch* batch, uint32_t slot,
                                             uint64_t insert_size_max, uint32_t winstride, uint32_t max_reads,
                                             char* header_buf, uint64_t header_cap, uint64_t* header_off) {
    if (!r || !batch) return mcb200_set_error(MCB200_EINVAL, "null argument");
    if (winstride == 0) return mcb200_set_error(MCB200_EINVAL, "winstride must be > 0");
    int64_t added = 0;
    uint64_t hpos = 0;
    if (header_off) header_off[0] = 0;
    std::string err;
    while (uint64_t(added) < max_reads) {
        if (!r->pending) {
            if (!next_pair(r, err)) {
                if (!err.empty()) return mcb200_set_error(MCB200_EIO, err.c_str());
                break;
            }
        }
        r->pending = true;
        if (header_buf && hpos + r->r1.hlen > header_cap) {                     // caller's header buffer is full
            if (added == 0) return mcb200_set_error(MCB200_EINVAL, "header buffer smaller than one header");
            break;
        }
        // make_candidate_generation_rules (candidate_structs.hpp:134-151)
        const uint32_t mw = uint32_t(2 + std::max<uint64_t>(r->r1.len + r->r2.len, insert_size_max) / winstride);
        const int rc = mcb200_batch_add_read(batch, slot, r->r1.seq, r->r1.len, r->r2.seq, r->r2.len, mw);
        if (rc < 0) return rc;
        if (rc == 0) break;                                                     // slot full: the pair stays pending
        r->pending = false;
        if (header_buf) { memcpy(header_buf + hpos, r->r1.hdr, r->r1.hlen); hpos += r->r1.hlen; }
        ++added;
        if (header_off) header_off[added] = hpos;
    }
    return added;
}
