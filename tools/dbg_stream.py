import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fuzzgen
from oracle import oracle
from brotli_rs_b200 import BatchDecoder, BroError, Decompressor
from test_gpu_parity import corpus_files
corpus = [c for _, c, _ in corpus_files()]
rng = np.random.default_rng(17)
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
for k, m in enumerate(fuzzgen.mutations(corpus, seed=21, count=250, max_len=20000)):
    chunk = int(rng.integers(1, 3000))
    if only >= 0 and k != only:
        continue
    st, out = oracle.decode(m)
    if "D" not in globals(): D = BatchDecoder(0)
    d = D
    got, err = b"", 0
    r = Decompressor(m, decoder=d, streaming=chunk)
    try:
        got = r.read()
    except BroError as e:
        err = e.status
    ok = err == st and (st != 0 or got == out)
    if not ok:
        print("MISMATCH", k, chunk, len(m), m[:24].hex(), "want", st, "got", err, d.last_error() if hasattr(d, "last_error") else "")
        break
    r.close()
print("done")
