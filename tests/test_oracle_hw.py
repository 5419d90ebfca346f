"""The adapter-search oracle (oracle/nr_oracle.c: nr_oracle_hw_search) against hand-written
known answers and its brute-force numpy twin.  edlib itself is not installed: parity unpinned,
the definition follows edlib's published semantics (see the C header comment)."""
import numpy as np


def test_known_answers(oracle):
    O = oracle
    # exact infix hit: end inclusive, like edlib
    assert O.hw_search("ACGT", "TTACGTTT", 1, False) == {
        "editDistance": 0, "n_locations": 1, "first": (2, 5), "last": (2, 5)}
    # one deleted base: two optimal ends (ACT / ACTT), same smallest start
    r = O.hw_search("ACGT", "TTACTTT", 1, False)
    assert r["editDistance"] == 1 and r["first"] == (2, 4) and r["last"] == (2, 5) and r["n_locations"] == 2
    # above k -> -1
    assert O.hw_search("ACGTACGT", "TTTTTTTTTT", 3, False)["editDistance"] == -1
    # N wildcard in the pattern matches any base only with the equalities (utils.py:15)
    assert O.hw_search("ACNNGT", "GGACTAGTGG", 0, True)["editDistance"] == 0
    assert O.hw_search("ACNNGT", "GGACTAGTGG", 1, False)["editDistance"] == -1
    # N in the text matches a pattern base with the equalities
    assert O.hw_search("ACGT", "TTANGTTT", 0, True)["editDistance"] == 0
    # repeated motif: first and last locations differ (the reference takes [-1] for 5' modes)
    r = O.hw_search("GATTACA", "CCGATTACACCGATTACACC", 2, False)
    assert r["editDistance"] == 0 and r["first"] == (2, 8) and r["last"] == (11, 17)
    # smallest start: a leading mismatch is preferred over a leading insertion
    r = O.hw_search("TACGT", "GGGACGTGG", 2, False)
    assert r["editDistance"] == 1 and r["first"] == (2, 6)
    # empty target
    assert O.hw_search("ACGT", "", 2, False)["editDistance"] == -1


def test_c_equals_numpy_twin(oracle):
    O = oracle
    rng = np.random.default_rng(5)
    pats = ["CGCTCTTCCGATCT" + "N" * 5 + "TTTCTT", "TCTCGGGAACGCTGAAGA", "ACGTN"]
    for it in range(120):
        pat = pats[it % 3]
        core = "".join(c if c != "N" else "ACGT"[rng.integers(0, 4)] for c in pat)
        c = list(core)
        for _ in range(int(rng.integers(0, 4))):
            j = int(rng.integers(0, len(c)))
            op = rng.random()
            if op < 0.33:
                c[j] = "ACGT"[rng.integers(0, 4)]
            elif op < 0.66:
                del c[j]
            else:
                c.insert(j, "ACGT"[rng.integers(0, 4)])
        alpha = "ACGTN" if it % 2 else "ACGT"
        t = ("".join(alpha[rng.integers(0, len(alpha))] for _ in range(int(rng.integers(0, 9)))) + "".join(c)
             + "".join(alpha[rng.integers(0, len(alpha))] for _ in range(int(rng.integers(0, 9)))))
        for wild in (True, False):
            k = min(4, len(pat) - 1)
            assert O.hw_search(pat, t, k, wild) == O.hw_search_numpy(pat, t, k, wild), (pat, t, wild)
