"""Reader fixtures for tests/test_host_decode.py (run in the build container, where /root/reference exists).

reader_cases.json holds, per case, an input file (base64) and what the UNMODIFIED reference does with it: whether
`colord compress-ont -q org` accepts it, the error line it prints if not, and — if it does — the file `colord decompress`
writes back (lossless mode, so the output shows exactly which headers, reads, '+' lines and qualities its reader took).
Usage: python tests/golden/make_reader_golden.py
"""
import base64
import gzip
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.abspath(os.path.join(HERE, "..", "..", "oracle", "_ref", "colord"))

R = [("r1 a=1", "ACGTNACGTTTGACCAGGATCCATTGACGTTAGCAAGTCCATGA", "I" * 44), ("r2", "TTTTGGGGCCAATTGGCCATATCGCGATTAGGCCTA", "!!!!((((**" + "5" * 26),
     ("r3/x", "GATTACAGATTACACCGGTTAACCGGTTAAGGCCTTAAGG", "5" * 40)]
L = "ACGTTGCAAGGCTTAACCGGATATCGCGTTAAGGCC"      # a read longer than k: very short reads make the reference itself crash


def fq(recs, eol="\n", plus=()):
    out = []
    for i, (h, s, q) in enumerate(recs):
        out += ["@" + h, s, "+" + (h if i in plus else ""), q]
    return (eol.join(out) + eol).encode()


CASES = {
    "plain": fq(R), "crlf": fq(R, "\r\n"), "cr_only": fq(R, "\r"), "plus_header": fq(R, plus=(0, 2)),
    "blank_lines_between_records": fq(R).replace(b"\n@r2", b"\n\n\n@r2") + b"\n\n",
    "blank_line_inside_record": fq(R).replace(b"\n+\n", b"\n\n+\n", 1),
    "blank_lines_after_first_record": ("@r0\n" + L + "\n+\n" + "I" * len(L) + "\n").encode() + b"\n\n" + fq(R),
    "quality_starts_with_at": fq([("r1", L, "@" + "I" * (len(L) - 1)), ("r2", L[::-1], "+@" + "I" * (len(L) - 2)), ("r3", L, "II@+" + "5" * (len(L) - 4))]),
    "header_only_at": fq([("", L, "I" * len(L)), ("r2", L[::-1], "I" * len(L))]),
    "header_with_tabs_and_spaces": fq([("r1\tx y  z", L, "I" * len(L)), ("r2", L[::-1], "I" * len(L))]),
    "later_header_without_at": ("@r1\n" + L + "\n+\n" + "I" * len(L) + "\nr2\n" + L[::-1] + "\n+\n" + "I" * len(L) + "\n").encode(),
    "plus_line_differs": b"@r1\nACGT\n+r2\nIIII\n", "plus_line_prefix_of_header": b"@r1x\nACGT\n+r1\nIIII\n",
    "lowercase": fq([("r", "ACgT", "IIII")]), "iupac": fq([("r", "ACRT", "IIII")]), "dot_in_read": fq([("r", "AC.T", "IIII")]),
    "all_n": fq([("r", "N" * 40, "I" * 40), ("s", L, "I" * len(L))]),
    "no_final_eol": fq(R)[:-1], "truncated_after_plus": b"@r1\nACGT\n+\nIIII\n@r2\nGGCC\n+\n", "truncated_after_read": b"@r1\nACGT\n+\nIIII\n@r2\nGGCC\n",
    "quality_shorter_than_read": b"@r1\nACGTAC\n+\nIIII\n@r2\nGG\n+\nIIII\n", "quality_length_mismatch_total": b"@r1\nACGTAC\n+\nIIII\n",
    "empty_file": b"", "only_newlines": b"\n\n\n", "unknown_format": b"ACGT\n", "starts_with_space": b" @r1\nACGT\n+\nIIII\n",
    "fasta_single_line": (">s1 d\n" + L + "\n>s2\n" + L[::-1] + "NN\n").encode(), "fasta_multi_line": (">s1 desc\n" + L + "\nACGT\r\n\r\nAC\n>s2\n" + "N" * 30 + "\n>s3\n" + L + "\nTT").encode(),
    "fasta_gt_inside_read_line": b">s1\nAC>GT\n", "fasta_header_then_header": b">s1\n>s2\nACGT\n", "fasta_lowercase": b">s1\nacgt\n",
    "fasta_no_final_eol": (">s1\n" + L).encode(), "fasta_blank_after_header": (">s1\n\n\n" + L + "\nGG\n>s2\n" + L[::-1] + "\n").encode(), "fasta_ends_after_header": (">s1\n" + L + "\n>s2\n").encode(),
}


def main():
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for name, data in CASES.items():
            fasta = data[:1] == b">"
            inp = os.path.join(d, "in.fa" if fasta else "in.fastq")
            open(inp, "wb").write(data)
            for f in ("a.colord", "out"):
                if os.path.exists(os.path.join(d, f)):
                    os.remove(os.path.join(d, f))
            c = subprocess.run([REF_BIN, "compress-ont", "-q", "org", "-t", "2", inp, "a.colord"], cwd=d, capture_output=True, text=True)
            text = c.stdout + c.stderr
            err = [l for l in text.split("\n") if l.startswith("Error") or "Only ACGTN" in l]
            entry = {"input_b64": base64.b64encode(data).decode(), "accepted": c.returncode == 0 and not err, "error": err[0] if err else "",
                     "crashed": c.returncode < 0}      # the reference segfaults / aborts on an assertion: no verdict of its reader
            if entry["accepted"]:
                dcmp = subprocess.run([REF_BIN, "decompress", "a.colord", "out"], cwd=d, capture_output=True, text=True)
                assert dcmp.returncode == 0, (name, dcmp.stderr)
                entry["output_b64"] = base64.b64encode(open(os.path.join(d, "out"), "rb").read()).decode()
            res[name] = entry
            print(f"{name:36s} {'ok ' if entry['accepted'] else 'CRASHED' if entry['crashed'] else 'REFUSED'} rc={c.returncode} {entry['error']}")
        # a gzipped copy of the plain case (the reference reads it through zlib)
        inp = os.path.join(d, "in.fastq.gz")
        with gzip.open(inp, "wb") as f:
            f.write(CASES["plain"])
        c = subprocess.run([REF_BIN, "compress-ont", "-q", "org", "-t", "2", inp, "g.colord"], cwd=d, capture_output=True, text=True)
        dcmp = subprocess.run([REF_BIN, "decompress", "g.colord", "gout"], cwd=d, capture_output=True, text=True)
        res["gzip"] = {"input_b64": base64.b64encode(open(inp, "rb").read()).decode(), "accepted": c.returncode == 0, "error": "",
                       "output_b64": base64.b64encode(open(os.path.join(d, "gout"), "rb").read()).decode()}
    json.dump(res, open(os.path.join(HERE, "reader_cases.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
