"""FASTQ texts shared by the ingest tests (oracle vs reference goldens, host emulation of the
kernels, GPU parity): hand-made corner cases of the reference loader (ReadData.cpp:177-198) and
seeded random texts."""
import numpy as np

EDGE_TEXTS = [
    b"",
    b"\n",
    b"x",
    b"@a\nACGT\n+\nIIII\n",
    b"@a\nACGT\n+\nIIII",                       # no trailing newline
    b"@a\nACGT\n+\nIIII\n@b\nGG",               # ends inside the read line
    b"@a\nACGT\n+\nIIII\n@b\n",                 # header only, terminated -> empty read
    b"@a\nACGT\n+\nIIII\n@b",                   # ends inside the header -> header bytes are the read
    b"@a\n\n+\n\n@b\nAC\r\n+\nII\n",            # empty read, CRLF kept as a base
    b"\n\n\n\n\n",
    b"@a\nACNNacgt\n+\nIIII\n\n",
    b"@h\nAC\n+",
    b"@h\nAC\n+\n",
    b"@h\nAC\n+\nII\n\n\n\n\n@x",
    b"@r\n" + b"ACGT" * 300 + b"\n+\n" + b"I" * 1200 + b"\n",
    b"@r\n" + b"G" * 15 + b"\n+\n\n@s\n" + b"T" * 17 + b"\n+\n\n@t\nC\n+\n\n",   # words straddling reads
    b"@q\nAC\x00GT\xffAC\n+\nIIIIIIII\n",       # any byte goes through the two bit tests
]


def random_fastq(rng, n_reads, mean_len, crlf=False, end="\n", alphabet=b"ACGTNacgt", short_frac=0.2):
    """Well-formed records with '@'/'+' inside quality strings; `end` = what follows the last quality line."""
    alpha = np.frombuffer(alphabet, np.uint8)
    out = []
    for i in range(n_reads):
        if rng.random() < short_frac:
            L = int(rng.integers(0, 40))
        else:
            L = int(rng.gamma(2.0, mean_len / 2.0))
        seq = rng.choice(alpha, size=L).tobytes()
        qual = rng.choice(np.frombuffer(b"!+@IJ#5", np.uint8), size=L).tobytes()
        eol = b"\r\n" if crlf else b"\n"
        out.append(b"@read" + str(i).encode() + b" x=" + b"y" * int(rng.integers(0, 60)) + eol + seq + eol + b"+" + eol + qual)
        out.append(eol if i + 1 < n_reads else end.encode() if isinstance(end, str) else end)
    return b"".join(out)


def expected_packed(bases, total_pad_words=0):
    """u32 words of the continuous 2-bit stream (first base most significant), as the device stores it."""
    b = np.ascontiguousarray(bases, dtype=np.uint8)
    code = ((b & 2) | ((b & 4) >> 2)).astype(np.uint32)
    nwords = (code.size + 15) // 16
    pad = np.zeros(nwords * 16, dtype=np.uint32)
    pad[:code.size] = code
    sh = (30 - 2 * np.arange(16)).astype(np.uint32)
    return (pad.reshape(nwords, 16) << sh).sum(axis=1).astype(np.uint32) if nwords else np.zeros(0, np.uint32)
