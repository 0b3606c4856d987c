"""Generates tests/golden/urban_sample.npz from the reference's sample urban network (run where /root/reference is
mounted):  python tests/golden/make_urban_sample.py

inputs : sampleData/urban-network.stn (UTM / MGA zone 55, orthometric heights), urban-network.msr (149 stations;
         types A B G H K L M S V Y Z, 1182 measurement rows, some flagged as ignored), urban-network.geo (geoid
         separations and deflections exported by dnageoid for these stations)
golden : sampleData/urban.phased.adj.expected — `dnaimport -> dnageoid --convert-stn-hts -> dnasegment -> dnaadjust
         urban --output-adj-msr --phased` (run-urban-network.sh).  A phased adjustment is rigorous, so its numbers are
         those of the simultaneous solution the oracle computes.

The .npz holds the binary station / measurement records (after the import, geoid and height-conversion steps, as
dnaadjust receives them) and the solution block, adjusted-measurement table and adjusted-coordinate table of the
expected file (angles in radians, angular corrections / standard deviations in seconds, as printed)."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import dna_ascii  # noqa: E402

SAMPLE = "/root/reference/sampleData"
ANGULAR = "ABKVZ"


def _dms(tokens):
    sign = -1.0 if tokens[0].startswith("-") else 1.0
    return sign * np.radians(abs(float(tokens[0])) + float(tokens[1]) / 60.0 + float(tokens[2]) / 3600.0)


def parse_expected(path):
    text = open(path).read()

    def grab(label):
        return float(re.search(r"^" + re.escape(label) + r"\s+(\S+)", text, re.M).group(1))
    sol = dict(unknowns=grab("Number of unknown parameters"), measurements=grab("Number of measurements"),
               dof=grab("Degrees of freedom"), chi_squared=grab("Chi squared"), sigma_zero=grab("Rigorous Sigma Zero"),
               pelzer=grab("Global (Pelzer) Reliability"), outliers=float(re.search(r"\((\d+) potential outliers\)", text).group(1)))
    body = text.split("Adjusted Measurements")[1].split("Adjusted Coordinates")[0]
    keys, rows = [], []
    for line in body.splitlines():
        if len(line) < 70 or line[0] not in "ABGHKLMSVYZ" or line[1] != " " or line.startswith("M Station"):
            continue
        t, f = line[0], line[62:].split()
        comp = ""
        if t in "GY":
            comp, f = f[0], f[1:]
        if t in ANGULAR or (t == "Y" and comp in "PL"):
            vals = [_dms(f[0:3]), _dms(f[3:6])] + [float(x) for x in f[6:13]]
        else:
            vals = [float(x) for x in f[0:9]]
        keys.append(" ".join([t, line[2:22].strip(), line[22:42].strip(), line[42:62].strip(), comp]).strip())
        rows.append(vals)
    stn_names, stn_rows = [], []
    for line in text.split("Adjusted Coordinates")[1].splitlines():
        f = line.split()
        if len(f) >= 12 and re.fullmatch(r"[CF]{3}", f[1]):
            stn_names.append(f[0])
            stn_rows.append([float(x) for x in f[2:12]])
    return sol, keys, np.array(rows), stn_names, np.array(stn_rows)


def main():
    stn = dna_ascii.read_stations(os.path.join(SAMPLE, "urban-network.stn"))
    dna_ascii.apply_geoid(stn, os.path.join(SAMPLE, "urban-network.geo"), convert_heights=True)
    msr = dna_ascii.read_measurements(os.path.join(SAMPLE, "urban-network.msr"), stn, reftran=False)
    sol, keys, rows, stn_names, stn_rows = parse_expected(os.path.join(SAMPLE, "urban.phased.adj.expected"))
    assert len(rows) == 1182 and len(stn_rows) == 149, (len(rows), len(stn_rows))
    # the exported geoid file carries N to a millimetre; the expected table prints both heights of every station to a tenth
    # of that, so the separation the run used is recovered as h(Ellipse) - H(Ortho) (an input of the run: the geoid model)
    where = {n.decode(): i for i, n in enumerate(stn["stationName"])}
    for n, row in zip(stn_names, stn_rows):
        i = where[n]
        stn["geoidSep"][i] = row[3] - row[2]
        stn["currentHeight"][i] = stn["initialHeight"][i] + float(stn["geoidSep"][i])
    out = os.path.join(ROOT, "tests", "golden", "urban_sample.npz")
    np.savez_compressed(out, stn=stn, msr=msr, solution_keys=np.array(sorted(sol)), solution=np.array([sol[k] for k in sorted(sol)]),
                        msr_keys=np.array(keys), msr_columns=np.array(["measured", "adjusted", "correction", "meas_sd", "adj_sd",
                                                                       "corr_sd", "nstat", "pelzer", "pre_adj_corr"]),
                        msr_rows=rows, stn_names=np.array(stn_names),
                        stn_columns=np.array(["lat_dms", "lon_dms", "H", "h", "X", "Y", "Z", "sd_e", "sd_n", "sd_up"]), stn_rows=stn_rows)
    print("wrote", out, os.path.getsize(out), "bytes;", sol)


if __name__ == "__main__":
    main()
