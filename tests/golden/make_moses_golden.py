"""Generates tests/golden/moses_golden.json: a corpus of strings with the token lists / detokenised strings the UNMODIFIED reference
tokenizer (/root/reference/mosestokenizer.cpp, built by oracle/Makefile into oracle/_ref/libmoses_ref.so) produces for them.
Run in the builder container (needs /root/reference):  python tests/golden/make_moses_golden.py"""
import ctypes as C
import json
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def corpus():
    base = [
        "Hello World!",
        "This ain't funny. It's actually hillarious, yet double Ls. | [] < > [ ] & You're gonna shake it off? Don't?",
        "Smith et al. The ... approx. two ... Fig. 2 and etc.",
        "disease. it spreads",
        "The patient (a 45-year-old male) received 5,300 mg of acetaminophen, i.e. too much.",
        "Dr. Smith and Mrs. Jones met Prof. X at St. Mary's Hosp. on Jan. 5, 1990.",
        "COVID-19 is caused by SARS-CoV-2.",
        "BRCA1/2 mutations increase risk by ~50%; p < 0.05, n = 1,234.",
        "What?! No... really?? Yes....",
        "He said, \"it's the '90s\", didn't he?",
        "e.g. the U.S.A. and the U.K. are abbreviations.",
        "Temperatures of 37.5°C – 40°C were recorded; TNF-α and IL-1β rose.",
        "naïve café résumé — “quoted” text… and ‘single’ quotes",
        "cost: $5, €10 or £3.50 (approx.)",
        "a,b,,c, d ,e 1,2 3, 4 ,5 6,",
        "the end.'",
        "trailing dots.. and more... and.... five.....",
        "x--y -- z - w a-b-c-d",
        "Nos. 1-3, No. 5, Art. 7 and pp. 10-12.",
        "O'Neil's dog's toys' colours; rock 'n' roll; 'tis",
        "[citation needed] <tag> a|b & c",
        "   leading and   trailing   spaces  ",
        "tab\tand\nnewline\r\nand \x07 bell",
        "",
        ".",
        "...",
        "'",
        "a.",
        "A. B. C.",
        "Mr.",
        "trastuzumab",
        "The 5'-UTR and 3' end; 1990's, 80's",
        "mid-1990s well-known state-of-the-art 10-fold",
        "www.example.com/path?a=1&b=2 user@mail.org",
        "pH 7.4 ± 0.2 (mean ± s.d.)",
    ]
    rng = random.Random(20260101)
    words = ["disease", "the", "patients", "Dr.", "et", "al.", "Fig.", "vs.", "i.e.", "e.g.", "approx.", "No.", "5,300", "3.14", "1,2,3", "it's", "don't",
             "O'Brien", "(", ")", "[", "]", "\"", "'", ",", ".", "...", "..", "?", "!", ";", ":", "-", "--", "a-b", "COVID-19", "α-helix", "β2", "naïve",
             "U.S.", "p<0.05", "&", "|", "<", ">", "50%", "$5", "€9", "“q”", "‘s’", "end.", "Mr.", "Mrs.", "St.", "Jan.", "pp.", "Art.", "Rs.", "Sept.", "x.y.z.",
             "1990's", "'90s", "can't.", "won't,", "is,", "2,", ",3", "T.", "rock'n'roll", "5'", "...and", "and...", "a.b", "3'-end"]
    out = list(base)
    for _ in range(300):
        k = rng.randint(1, 14)
        parts = [rng.choice(words) for _ in range(k)]
        sep = [rng.choice([" ", " ", " ", "", "  "]) for _ in range(k)]
        out.append("".join(p + s for p, s in zip(parts, sep)))
    return out


def tokenize_file(path_in, path_out):
    """helper for tests/test_text.py (runs in its own process: the reference library must not share a process with ours, both
    export moses_tokenize): JSON list of strings in -> JSON list of token lists (None where the reference throws) out"""
    texts = json.load(open(path_in))
    os.chdir("/root/reference/examples")
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmoses_ref.so"))
    L.refmoses_tokenize.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(1 << 16)
    out = []
    for t in texts:
        n = L.refmoses_tokenize(t.encode("utf-8", errors="surrogateescape"), b"en", buf, len(buf))
        out.append(None if n == -2 else (buf.value.decode("utf-8", errors="surrogateescape").split("\n") if n else []))
    json.dump(out, open(path_out, "w"))


def main():
    here = os.getcwd()
    os.chdir("/root/reference/examples")          # ../data/perluniprops must resolve while the library's static initialisers run and on every call
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmoses_ref.so"))
    L.refmoses_tokenize.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    L.refmoses_detokenize.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(1 << 16)
    cases = []
    for text in corpus():
        n = L.refmoses_tokenize(text.encode("utf-8"), b"en", buf, len(buf))
        if n == -2:                                # the reference throws on this input (see oracle/ref_moses_shim.cpp): nothing to pin
            cases.append({"text": text, "reference_throws": True})
            continue
        assert n >= 0
        toks = buf.value.decode("utf-8", errors="surrogateescape").split("\n") if n else []
        m = L.refmoses_detokenize("\n".join(toks).encode("utf-8", errors="surrogateescape"), b"en", buf, len(buf))
        assert m >= 0
        cases.append({"text": text, "tokens": toks, "detok": buf.value.decode("utf-8", errors="surrogateescape")})
    os.chdir(here)
    json.dump({"source": "reference moses_tokenize / moses_detokenize (lang en), /root/reference @ e07668a", "cases": cases},
              open(os.path.join(ROOT, "tests", "golden", "moses_golden.json"), "w"), ensure_ascii=True, indent=0)
    print(len(cases), "cases")


if __name__ == "__main__":
    import sys
    if len(sys.argv) == 4 and sys.argv[1] == "--tokenize":
        tokenize_file(sys.argv[2], sys.argv[3])
    else:
        main()
