"""Generates tests/golden/urban_mt_sample.npz from the reference's sample urban network (run where /root/reference is
mounted):  python tests/golden/make_urban_mt_sample.py

inputs : sampleData/urban-network.stn / .msr (GDA94) and urban-network.geo, as for urban_sample.npz
golden : sampleData/urban_mt.phased-mt.adj.expected — the reference's CI chain `dnaimport -n urban_mt` ->
         `dnareftran urban_mt -r gda2020` -> `dnageoid` -> `dnasegment` -> `dnaadjust urban_mt ... --free-stn-sd 4.0
         --fixed-stn-sd 0.000001 --max-iterations 20 ...` -> `dnaadjust urban_mt --output-adj-msr --multi`
         (CMakeLists.txt:1077-1083, compared by dnadiff at 0.01, :1191).  The expected file is the SECOND adjustment: it
         starts from the binary files the first one reduced and updated, so it pins the re-adjustment path
         (metadata `reduced`, ADJ:296, 3913-3935) and, through the reference-frame step, GNSS data moved GDA94 -> GDA2020.

The .npz holds the records as the first dnaadjust receives them and the tables of the expected file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import dna_ascii  # noqa: E402
from tests.golden.make_urban_sample import parse_expected  # noqa: E402

SAMPLE = "/root/reference/sampleData"


def main():
    stn = dna_ascii.read_stations(os.path.join(SAMPLE, "urban-network.stn"))
    dna_ascii.reftran_stations(stn, "GDA94")
    dna_ascii.apply_geoid(stn, os.path.join(SAMPLE, "urban-network.geo"), convert_heights=True)
    msr = dna_ascii.read_measurements(os.path.join(SAMPLE, "urban-network.msr"), stn, reftran=True)
    sol, keys, rows, stn_names, stn_rows = parse_expected(os.path.join(SAMPLE, "urban_mt.phased-mt.adj.expected"))
    assert len(rows) == 1182 and len(stn_rows) == 149, (len(rows), len(stn_rows))
    # the exported geoid file carries N to a millimetre; the expected table prints both heights of every station to a tenth
    # of that, so the separation the run used is recovered as h(Ellipse) - H(Ortho) (an input of the run: the geoid model)
    where = {n.decode(): i for i, n in enumerate(stn["stationName"])}
    for n, row in zip(stn_names, stn_rows):
        i = where[n]
        stn["geoidSep"][i] = row[3] - row[2]
        stn["currentHeight"][i] = stn["initialHeight"][i] + float(stn["geoidSep"][i])
    out = os.path.join(ROOT, "tests", "golden", "urban_mt_sample.npz")
    np.savez_compressed(out, stn=stn, msr=msr, solution_keys=np.array(sorted(sol)), solution=np.array([sol[k] for k in sorted(sol)]),
                        msr_keys=np.array(keys), msr_rows=rows, stn_names=np.array(stn_names), stn_rows=stn_rows)
    print("wrote", out, os.path.getsize(out), "bytes;", sol)


if __name__ == "__main__":
    main()
