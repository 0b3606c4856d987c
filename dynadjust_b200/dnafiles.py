"""DynAdjust binary / ASCII files for the adjust step (Python side; the C++ twin is csrc/host/dna_files.hpp).

.bst / .bms: 60-byte text header + metadata + raw record dump (include/io/dynadjust_file.cpp:67-116, 190-283);
.seg: dnasegment's ASCII block lists (include/io/seg_file.cpp:489-721)."""
import datetime
import struct

import numpy as np

from .records import MSR_DTYPE, STN_DTYPE

FIELD = 10


def _header(version="1.2", app="DNA10400"):
    date = datetime.date.today().isoformat()
    return (b"VERSION   " + version.rjust(FIELD).encode() + b"CREATED ON" + date.rjust(FIELD).encode()
            + b"CREATED BY" + app[:FIELD].rjust(FIELD).encode())


def _metadata(count, reduced=False, modified_by="import", epsg="7843", epoch="01.01.2020", obs_epoch="", reftran=False, geoid=False):
    def fixed(s, n):
        return s.encode()[:n - 1].ljust(n, b"\0")
    out = struct.pack("<Q?", count, reduced) + fixed(modified_by, 20) + fixed(epsg, 7) + fixed(epoch, 12) + fixed(obs_epoch, 12)
    out += struct.pack("<??", reftran, geoid)
    out += struct.pack("<Q", 0)   # input files
    out += struct.pack("<Q", 0)   # source files
    return out


def write_binary(path, records, **meta):
    with open(path, "wb") as f:
        f.write(_header())
        f.write(_metadata(len(records), **meta))
        f.write(np.ascontiguousarray(records).tobytes())


def read_binary(path, dtype):
    with open(path, "rb") as f:
        data = f.read()
    pos = 60
    count, reduced = struct.unpack_from("<Q?", data, pos)
    pos += 9 + 20 + 7 + 12 + 12 + 2
    (nin,) = struct.unpack_from("<Q", data, pos)
    pos += 8 + nin * (256 + 7 + 12 + 12 + 4)
    (nsrc,) = struct.unpack_from("<Q", data, pos)
    pos += 8 + nsrc * 256
    recs = np.frombuffer(data, dtype=dtype, count=count, offset=pos).copy()
    return recs, dict(reduced=bool(reduced), version=data[10:20].decode().strip())


def write_bst(path, stn, **meta):
    assert stn.dtype == STN_DTYPE
    write_binary(path, stn, **meta)


def write_bms(path, msr, **meta):
    assert msr.dtype == MSR_DTYPE
    write_binary(path, msr, **meta)


def write_seg(path, isl, jsl, cml, bst="", bms=""):
    """Chain segmentation in dnasegment's layout (what the reference's SegFile::LoadSegFile parses)."""
    dash = "-" * 80
    L = [dash, "DYNADJUST SEGMENTATION OUTPUT FILE", "",
         f"{'Version:':<35}1.2.9", f"{'Build:':<35}-", f"{'File created:':<35}-", f"{'File name:':<35}{path}", "",
         f"{'Command line arguments:':<35}-", "", f"{'Stations file:':<35}{bst}", f"{'Measurements file:':<35}{bms}", "",
         f"{'Minimum inner stations':<35}150", f"{'Block size threshold':<35}150", f"{'Starting station(s)':<35}-", dash, "",
         "SEGMENTATION SUMMARY", "", f"{'No. blocks produced':<35}{len(isl)}", dash,
         f"{'Block':<14}{'Network ID':<14}{'Junction stns':<16}{'Inner stns':<16}{'Measurements':<16}{'Total stns':<16}"]
    for b in range(len(isl)):
        L.append(f"{b + 1:<14}{0:<14}{len(jsl[b]):<16}{len(isl[b]):<16}{len(cml[b]):<16}{len(isl[b]) + len(jsl[b]):<16}")
    L += [dash, "", "INDIVIDUAL BLOCK DATA", dash]
    for b in range(len(isl)):
        L += ["", f"Block {b + 1}", dash, f"{'Junction stns:':<35}{len(jsl[b])}", f"{'Inner stns:':<35}{len(isl[b])}",
              f"{'Measurements:':<35}{len(cml[b])}", f"{'Total stns:':<35}{len(isl[b]) + len(jsl[b])}", "",
              f"{'Inner stns':<16}{'Junction stns':<16}{'Measurements':<16}", dash]
        rows = max(len(isl[b]), len(jsl[b]), len(cml[b]))
        for r in range(rows):
            a = str(isl[b][r]) if r < len(isl[b]) else ""
            c = str(jsl[b][r]) if r < len(jsl[b]) else ""
            d = str(cml[b][r]) if r < len(cml[b]) else ""
            L.append(f"{a:<16}{c:<16}{d:<16}")
        L.append(dash)
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")


def write_asl(path, assoc_msr_count, aml_index, validity):
    """<net>.asl (include/io/asl_file.cpp:80-93): header, u64 count, then {u32 assocMsrCount; u32 amlStnIndex; u16 validity}."""
    with open(path, "wb") as f:
        f.write(_header())
        f.write(struct.pack("<Q", len(validity)))
        for a, i, v in zip(assoc_msr_count, aml_index, validity):
            f.write(struct.pack("<IIH", int(a), int(i), int(v)))


def write_map(path, names):
    """<net>.map (include/io/map_file.cpp:44-98): header, u32 count, then {char name[31]; u32 bstIndex} sorted by name."""
    order = sorted(range(len(names)), key=lambda i: names[i])
    with open(path, "wb") as f:
        f.write(_header())
        f.write(struct.pack("<I", len(names)))
        for i in order:
            nm = names[i] if isinstance(names[i], bytes) else names[i].encode()
            f.write(nm[:30].ljust(31, b"\0"))
            f.write(struct.pack("<I", i))
