"""NumPy views of the DynAdjust binary records (.bst / .bms payloads).

Layout mirrors ``include/dna_records.h`` (bit-compatible with the reference's
``msr_t`` — dynadjust/include/measurement_types/dnameasurement.hpp:133-194, 208 bytes —
and ``stn_t`` — dynadjust/include/config/dnatypes-structs.hpp:270-323, 352 bytes).
"""
import numpy as np

MSR_DTYPE = np.dtype({
    "names": ["measType", "measStart", "measurementStations", "epsgCode", "epoch", "observation_epoch",
              "coordType", "ignore", "station1", "station2", "station3", "vectorCount1", "vectorCount2",
              "clusterID", "fileOrder", "sourceFileIndex", "term1", "term2", "term3", "term4",
              "scale1", "scale2", "scale3", "scale4", "measAdj", "measCorr", "measAdjPrec", "residualPrec",
              "NStat", "TStat", "PelzerRel", "preAdjCorr", "preAdjMeas"],
    "formats": ["S1", "i1", "i1", "S7", "S12", "S12", "S4", "u1", "<u4", "<u4", "<u4", "<u4", "<u4",
                "<u4", "<u4", "<u4", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8",
                "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8"],
    "offsets": [0, 1, 2, 3, 10, 22, 34, 38, 40, 44, 48, 52, 56, 60, 64, 68, 72, 80, 88, 96,
                104, 112, 120, 128, 136, 144, 152, 160, 168, 176, 184, 192, 200],
    "itemsize": 208,
})

STN_DTYPE = np.dtype({
    "names": ["stationName", "stationNameOrig", "stationConst", "stationType", "suppliedStationType",
              "initialLatitude", "currentLatitude", "initialLongitude", "currentLongitude",
              "initialHeight", "currentHeight", "suppliedHeightRefFrame", "geoidSep", "geoidSepUnc",
              "meridianDef", "verticalDef", "zone", "description", "fileOrder", "nameOrder", "clusterID",
              "unusedStation", "epsgCode", "epoch", "observation_epoch", "plate"],
    "formats": ["S31", "S40", "S4", "S4", "<u2", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "<u2", "<f4", "<f4",
                "<f8", "<f8", "<i2", "S129", "<u4", "<u4", "<u4", "<u2", "S7", "S12", "S12", "S3"],
    "offsets": [0, 31, 71, 75, 80, 88, 96, 104, 112, 120, 128, 136, 140, 144,
                152, 160, 168, 170, 300, 304, 308, 312, 314, 321, 333, 345],
    "itemsize": 352,
})

assert MSR_DTYPE.itemsize == 208 and STN_DTYPE.itemsize == 352

# suppliedStationType (dnatypes-basic.hpp:127-135)
XYZ_TYPE, LLh_TYPE, LLH_TYPE, UTM_TYPE = 0, 1, 2, 3

# GRS80 / GDA2020 (dnaconsts.hpp:47-48, dnadatumprojectionparam.hpp:38-39)
GRS80_A = 6378137.0
GRS80_INVF = 298.257222101


def new_msr(count):
    """Zeroed measurement records with the reference constructor's defaults (scale1..4 = 1, epsg 7843)."""
    m = np.zeros(count, dtype=MSR_DTYPE)
    m["scale1"] = m["scale2"] = m["scale3"] = m["scale4"] = 1.0
    m["epsgCode"] = b"7843"
    m["measurementStations"] = 1
    return m


def new_stn(count):
    s = np.zeros(count, dtype=STN_DTYPE)
    s["suppliedStationType"] = LLH_TYPE
    s["suppliedHeightRefFrame"] = 0
    s["epsgCode"] = b"7843"
    return s
