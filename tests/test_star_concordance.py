"""The STAR concordance harness (tools/star_concordance.py), exercised without STAR: with no
binary on PATH it must say so; with a stub `STAR` on PATH (a script that answers the two
invocations of scripts/barcode_ref.sh / barcode_align.sh from the oracle) every code path --
padded FASTA, genomeGenerate argv, align argv, the rename of barcode_align.sh:41, SAM parse,
comparison -- runs and reports full concordance."""
import os
import stat
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

STUB = textwrap.dedent('''\
    #!%(py)s
    """stub STAR: answers genomeGenerate and the barcode alignment from the CPU oracle"""
    import gzip, os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np
    from oracle import oracle as O
    a = sys.argv[1:]
    opt = {}
    i = 0
    while i < len(a):
        j = i + 1
        while j < len(a) and not a[j].startswith("--"):
            j += 1
        opt[a[i]] = a[i + 1:j]
        i = j
    if "--version" in opt:
        print("2.7.9a-stub"); sys.exit(0)
    if opt.get("--runMode") == ["genomeGenerate"]:
        assert opt["--genomeSAindexNbases"] == ["6"] and opt["--genomeChrBinNbits"] == ["7"]
        gd = opt["--genomeDir"][0]
        os.makedirs(gd, exist_ok=True)
        open(os.path.join(gd, "fasta_path.txt"), "w").write(opt["--genomeFastaFiles"][0])
        sys.exit(0)
    # alignment: every flag of barcode_align.sh must be there
    for f, v in (("--alignEndsType", ["EndToEnd"]), ("--outFilterMultimapNmax", ["1"]),
                 ("--outSAMattributes", ["AS", "nM", "MD"]), ("--scoreInsBase", ["-1"]),
                 ("--scoreDelBase", ["-1"]), ("--outFilterScoreMinOverLread", ["0"]),
                 ("--readFilesCommand", ["zcat"]), ("--seedSearchStartLmax", ["4"])):
        assert opt.get(f) == v, (f, opt.get(f))
    fa = open(os.path.join(opt["--genomeDir"][0], "fasta_path.txt")).read()
    names, recs = [], []
    for ln in open(fa):
        ln = ln.strip()
        (names if ln.startswith(">") else recs).append(ln.lstrip(">"))
    left = len(recs[0]) - len(recs[0].lstrip("N"))
    right = len(recs[0]) - len(recs[0].rstrip("N"))
    cores = [r[left:len(r) - right] for r in recs]
    qn, qs = [], []
    for ln in gzip.open(opt["--readFilesIn"][0], "rt"):
        ln = ln.strip()
        (qn if ln.startswith(">") else qs).append(ln.lstrip(">").split(" ")[0])
    wlc, _ = O.encode_many(cores, len(cores[0]))
    cc, cl = O.encode_many(qs, 64)
    r = O.match(wlc, left, right, cc, cl)
    with open(opt["--outFileNamePrefix"][0] + "Aligned.out.sam", "w") as f:
        for n in names:
            f.write("@SQ\\tSN:%%s\\tLN:%%d\\n" %% (n, len(recs[0])))
        for k in range(len(qn)):
            if r["n_best"][k] != 1:
                continue
            fl = 16 if r["strand"][k] else 0
            f.write("\\t".join([qn[k], str(fl), names[r["best_idx"][k]], "1", "255", "%%dM" %% len(qs[k]),
                               "*", "0", "0", qs[k], "*", "NH:i:1", "HI:i:1",
                               "AS:i:%%d" %% r["best_score"][k], "nM:i:0"]) + "\\n")
''')


def test_reports_absence_without_star(monkeypatch, tmp_path):
    import star_concordance as SC
    monkeypatch.setenv("PATH", str(tmp_path))          # nothing there
    r = SC.run()
    assert r["status"] == "STAR absent — concordance not measured"
    assert r["star_binary"] == "absent"


def test_stub_star_exercises_the_whole_harness(monkeypatch, tmp_path, oracle):
    import star_concordance as SC
    stub = tmp_path / "STAR"
    stub.write_text(STUB % {"py": sys.executable, "root": ROOT})
    stub.chmod(stub.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    monkeypatch.delenv("NANORANGER_REF", raising=False)
    r = SC.run(fixtures=["slideseq"], threads=2)
    assert r["status"] == "measured", r
    assert r["star_version"].startswith("2.7.9a-stub")
    f = r["fixtures"]["slideseq"]
    assert f["invocation"] == "restated argv"
    assert f["n"] == 1511 and f["star_records"] > 1000
    # the stub answers from the oracle: everything it reports is identical, nothing differs
    assert f["identical_barcode_as"] == f["star_records"]
    assert f["star_only"] == 0 and f["oracle_only"] == 0 and f["star_differs_from_unique_oracle_pair"] == 0
    assert f["index_build_s"] > 0 and f["align_s"] > 0


def test_argv_tables_restate_every_script_flag():
    """when the reference checkout is present (build container only) the restated argv must equal
    the scripts' flag for flag; on boxes without it the test is skipped."""
    import re
    import star_concordance as SC
    ref = os.environ.get("NANORANGER_REF", "/root/reference")
    path = os.path.join(ref, "scripts", "barcode_align.sh")
    if not os.path.isfile(path):
        pytest.skip("reference checkout not present")
    for script, table in (("barcode_align.sh", SC.ALIGN_FLAGS), ("barcode_ref.sh", SC.GENOME_GENERATE_FLAGS)):
        txt = open(os.path.join(ref, "scripts", script)).read()
        flags = re.findall(r"^(--\S+)\s+(.*?)\s*\\?$", txt, flags=re.M)
        assert [f for f, _ in flags] == [row[0] for row in table], script
        for (f, v), row in zip(flags, table):
            if "$" not in v:
                assert v.split() == list(row[1:]), (script, f)
