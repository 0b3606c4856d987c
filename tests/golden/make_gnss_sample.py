"""Generates tests/golden/gnss_sample.npz from the reference's sample GNSS network (run in the build container, where
/root/reference is mounted; the .npz travels to the GPU box):

    python tests/golden/make_gnss_sample.py

inputs : sampleData/gnss-network.stn, gnss-network.msr  (ASCII DNA files; read by tests/golden/dna_ascii.py, which
         emulates dnaimport + dnareftran for the G / X / Y types of this network)
golden : sampleData/gnss.simult.adj.expected — the reference's own expected output of
         `dnaimport -> dnageoid -> dnareftran -> dnaadjust gnss --output-adj-msr` (CMakeLists.txt:1027-1034), compared by
         its CI with dnadiff at tolerance 0.001.

The .npz holds the binary station / measurement records handed to the adjustment and the numbers printed in the
expected file: the solution block, the adjusted-measurement table and the adjusted-coordinate table."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import dna_ascii  # noqa: E402

SAMPLE = "/root/reference/sampleData"


def parse_expected(path):
    text = open(path).read()

    def grab(label):
        return float(re.search(r"^" + re.escape(label) + r"\s+(\S+)", text, re.M).group(1))
    sol = dict(unknowns=grab("Number of unknown parameters"), measurements=grab("Number of measurements"),
               dof=grab("Degrees of freedom"), chi_squared=grab("Chi squared"), sigma_zero=grab("Rigorous Sigma Zero"),
               pelzer=grab("Global (Pelzer) Reliability"), outliers=float(re.search(r"\((\d+) potential outliers\)", text).group(1)),
               iterations=float(len(re.findall(r"^ITERATION\s+\d+", text, re.M))))
    lo, hi = re.search(r"Chi-Square test \(95.0%\)\s+(\S+) < \S+ < (\S+)", text).groups()
    sol["chi_lower"], sol["chi_upper"] = float(lo), float(hi)
    body = text.split("Adjusted Measurements")[1].split("Adjusted Coordinates")[0]
    msr_rows, msr_keys = [], []
    for line in body.splitlines():
        f = line.split()
        if len(f) < 12 or f[0] not in ("G", "X", "Y") or f[1] == "Station":
            continue
        # type, station(s), component, then nine numbers (an outlier flag may follow)
        k = next(i for i in range(2, len(f)) if f[i] in ("X", "Y", "Z") and re.fullmatch(r"-?\d+\.\d+", f[i + 1]))
        msr_keys.append(" ".join([f[0]] + f[1:k] + [f[k]]))
        msr_rows.append([float(x) for x in f[k + 1:k + 10]])
    stn_rows, stn_names = [], []
    for line in text.split("Adjusted Coordinates")[1].splitlines():
        f = line.split()
        if len(f) >= 12 and re.fullmatch(r"[CF]{3}", f[1]):
            stn_names.append(f[0])
            stn_rows.append([float(x) for x in f[2:12]])
    return sol, msr_keys, np.array(msr_rows), stn_names, np.array(stn_rows)


def main():
    stn = dna_ascii.read_stations(os.path.join(SAMPLE, "gnss-network.stn"))
    msr = dna_ascii.read_measurements(os.path.join(SAMPLE, "gnss-network.msr"), stn)
    sol, msr_keys, msr_rows, stn_names, stn_rows = parse_expected(os.path.join(SAMPLE, "gnss.simult.adj.expected"))
    assert len(msr_rows) == 417 and len(stn_rows) == 43
    # dnaimport writes the binary station file sorted by name (fileOrder keeps the order of the input file): same here,
    # with the station indices of the measurement records following
    names = [n.decode() for n in stn["stationName"]]
    order = sorted(range(len(stn)), key=lambda i: names[i])
    new_index = np.empty(len(stn), np.uint32)
    new_index[order] = np.arange(len(stn), dtype=np.uint32)
    stn = stn[order].copy()
    msr["station1"] = new_index[msr["station1"]]
    two = msr["measType"] != b"Y"
    msr["station2"][two] = new_index[msr["station2"][two]]
    # dnageoid stores the geoid separation of every station; the grid file is not reproduced here, but the expected table
    # prints both heights, so N = h(Ellipse) - H(Ortho) is recovered from it (an input of the run, not a result: the
    # adjustment of this GNSS-only network does not use it)
    where = {n.decode(): i for i, n in enumerate(stn["stationName"])}
    for n, row in zip(stn_names, stn_rows):
        stn["geoidSep"][where[n]] = row[3] - row[2]
    out = os.path.join(ROOT, "tests", "golden", "gnss_sample.npz")
    np.savez_compressed(out, stn=stn, msr=msr, solution_keys=np.array(sorted(sol)), solution=np.array([sol[k] for k in sorted(sol)]),
                        msr_keys=np.array(msr_keys), msr_columns=np.array(["measured", "adjusted", "correction", "meas_sd", "adj_sd",
                                                                           "corr_sd", "nstat", "pelzer", "pre_adj_corr"]),
                        msr_rows=msr_rows, stn_names=np.array(stn_names),
                        stn_columns=np.array(["lat_dms", "lon_dms", "H", "h", "X", "Y", "Z", "sd_e", "sd_n", "sd_up"]), stn_rows=stn_rows)
    print("wrote", out, os.path.getsize(out), "bytes;", sol)


if __name__ == "__main__":
    main()
