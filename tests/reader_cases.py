"""Input files of the FASTA/FASTQ reader tests.  The expected records come from the reference's own
reader (oracle/make_reader_golden.py -> tests/golden/reader.json)."""
import numpy as np


def big_content(recipe):
    kind, n, seed = recipe
    rng = np.random.default_rng(seed)
    out = bytearray()
    if kind == "fasta_long_lines":          # single-line sequences far longer than the reference's 64 KB buffer
        for i in range(n):
            ln = int(rng.integers(1, 300_000))
            out += b">long%d some text\n" % i + rng.choice(list(b"ACGTN"), ln).astype(np.uint8).tobytes() + b"\n"
    elif kind == "fasta_wrapped":           # 60-column FASTA, empty lines, CRLF now and then
        for i in range(n):
            ln = int(rng.integers(0, 5_000))
            s = rng.choice(list(b"ACGTacgtN"), ln).astype(np.uint8).tobytes()
            eol = b"\r\n" if i % 7 == 3 else b"\n"
            out += b">r%d|x" % i + eol
            for p in range(0, ln, 60):
                out += s[p:p + 60] + eol
                if (i + p) % 997 == 0:
                    out += b"\n"
    elif kind == "fastq":                   # 4-line FASTQ whose quality lines often start with '@' or '+'
        for i in range(n):
            ln = int(rng.integers(1, 400))
            s = rng.choice(list(b"ACGT"), ln).astype(np.uint8).tobytes()
            q = rng.choice(list(b"@+>IIIIFFFF#"), ln).astype(np.uint8).tobytes()
            out += b"@read%d/1\n" % i + s + b"\n+\n" + q + b"\n"
    return bytes(out)


CASES = [
    dict(name="fasta_simple", files=[("raw", b">a\nACGT\n>b desc\nGGGTTT\n")]),
    dict(name="fasta_multiline", files=[("raw", b">a\nAC\nGT\n\nTT\n>b\n\n\nA\n>c\n>d\nNN\n")]),
    dict(name="fasta_crlf", files=[("raw", b">a x\r\nACGT\r\nGG\r\n>b\r\nTT\r\n")]),
    dict(name="fasta_no_final_newline", files=[("raw", b">a\nACGT\n>b\nGG")]),
    dict(name="fasta_header_only_at_end", files=[("raw", b">a\nACGT\n>b")]),
    dict(name="fasta_empty_header", files=[("raw", b">\nACGT\n>\n\n>x\nA\n")]),
    dict(name="fasta_at_in_data", files=[("raw", b">a\nAC\n@notaheader\nGT\n>b\nT\n")]),
    dict(name="fasta_plus_in_data", files=[("raw", b">a\nAC\n+\nGT\nTT\n>b\nT\n")]),
    dict(name="fasta_lower_iupac", files=[("raw", b">a\nacgtnNRYKMswbdhv-.\n")]),
    dict(name="fastq_simple", files=[("raw", b"@r1\nACGT\n+\nIIII\n@r2 x\nGG\n+r2\nII\n")]),
    dict(name="fastq_qual_starts_with_at", files=[("raw", b"@r1\nACGT\n+\n@III\n@r2\nGG\n+\n+I\n@r3\nT\n+\n>\n")]),
    dict(name="fastq_multiline_seq", files=[("raw", b"@r1\nAC\nGT\n+\nIIII\n@r2\nGG\n+\nII\n")]),
    dict(name="fastq_crlf", files=[("raw", b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGG\r\n+\r\nII\r\n")]),
    dict(name="fastq_truncated", files=[("raw", b"@r1\nACGT\n+\nIIII\n@r2\nGG\n+")]),
    dict(name="fastq_truncated2", files=[("raw", b"@r1\nACGT\n+\nIIII\n@r2\nGG")]),
    dict(name="fastq_extra_quality_line", files=[("raw", b"@r1\nACGT\n+\nII\nII\n@r2\nGG\n+\nII\n")]),
    dict(name="garbage_between_records", files=[("raw", b">a\nACGT\n+\nqual\ngarbage line\nmore\n>b\nTT\n")]),
    dict(name="malformed_start", files=[("raw", b"ACGT\n>a\nAC\n")]),
    dict(name="empty_file", files=[("raw", b"")]),
    dict(name="only_newlines_after_header", files=[("raw", b">a\n\n\n\n")]),
    dict(name="gz_fasta", gz=True, files=[("raw", b">a\nACGT\nAC\n>b\nGG\n")]),
    dict(name="gz_fastq", gz=True, files=[("raw", b"@r1\nACGT\n+\nIIII\n@r2\nGG\n+\nII\n")]),
    dict(name="pairfiles_equal", files=[("raw", b">a/1\nACGT\n>b/1\nGG\n"), ("raw", b">a/2\nTTTT\n>b/2\nCC\n")]),
    dict(name="pairfiles_second_shorter", files=[("raw", b">a/1\nACGT\n>b/1\nGG\n>c/1\nA\n"), ("raw", b">a/2\nTTTT\n")]),
    dict(name="pairfiles_first_shorter", files=[("raw", b">a/1\nACGT\n"), ("raw", b">a/2\nTTTT\n>b/2\nCC\n")]),
    dict(name="pairseq_even", pairseq=True, files=[("raw", b">a/1\nACGT\n>a/2\nTT\n>b/1\nGG\n>b/2\nCC\n")]),
    dict(name="pairseq_odd", pairseq=True, files=[("raw", b">a/1\nACGT\n>a/2\nTT\n>b/1\nGG\n")]),
    dict(name="pairseq_fastq", pairseq=True, files=[("raw", b"@a/1\nACGT\n+\nIIII\n@a/2\nTT\n+\nII\n")]),
    dict(name="big_fasta_long_lines", files=[("big", ("fasta_long_lines", 40, 1))]),
    dict(name="big_fasta_wrapped", files=[("big", ("fasta_wrapped", 3000, 2))]),
    dict(name="big_fastq", files=[("big", ("fastq", 20000, 3))]),
    dict(name="big_fastq_gz", gz=True, files=[("big", ("fastq", 5000, 4))]),
    dict(name="big_pairseq", pairseq=True, files=[("big", ("fasta_wrapped", 2001, 5))]),
]
