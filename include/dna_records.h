/*
 * dna_records.h — binary station / measurement records exchanged with the
 * DynAdjust tool chain (.bst / .bms payloads).
 *
 * These are layout mirrors, written from the field list and the x86-64 natural
 * alignment of the reference structs; they are bit-compatible with
 *   msr_t  — dynadjust/include/measurement_types/dnameasurement.hpp:133-194  (208 bytes)
 *   stn_t  — dynadjust/include/config/dnatypes-structs.hpp:270-323          (352 bytes)
 * Field widths: dynadjust/include/config/dnatypes-basic.hpp:66-76.
 * The static_asserts below pin every offset the adjustment path touches.
 */
#ifndef DNA_RECORDS_H_
#define DNA_RECORDS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    DNA_STN_NAME_WIDTH = 31,
    DNA_STN_NAME_ORIG_WIDTH = 40,
    DNA_STN_DESC_WIDTH = 129,
    DNA_STN_CONST_WIDTH = 4,
    DNA_STN_TYPE_WIDTH = 4,
    DNA_STN_EPSG_WIDTH = 7,
    DNA_STN_EPOCH_WIDTH = 12,
    DNA_STN_PLATE_WIDTH = 3
};

/* measStart values (dnatypes-basic.hpp:173-180) */
enum { DNA_X_MEAS = 0, DNA_Y_MEAS = 1, DNA_Z_MEAS = 2, DNA_X_COV = 3, DNA_Y_COV = 4, DNA_Z_COV = 5 };

/* suppliedStationType values (dnatypes-basic.hpp:127-135) */
enum { DNA_XYZ_TYPE = 0, DNA_LLh_TYPE = 1, DNA_LLH_TYPE = 2, DNA_UTM_TYPE = 3, DNA_ENU_TYPE = 4, DNA_AED_TYPE = 5 };

typedef struct dna_msr_t {
    char measType;              /* 'G','X','Y','D','S','L',... */
    char measStart;             /* DNA_X_MEAS .. DNA_Z_COV */
    char measurementStations;   /* 1, 2 or 3 */
    char epsgCode[DNA_STN_EPSG_WIDTH];
    char epoch[DNA_STN_EPOCH_WIDTH];
    char observation_epoch[DNA_STN_EPOCH_WIDTH];
    char coordType[4];
    uint8_t ignore;             /* bool */
    uint32_t station1;
    uint32_t station2;
    uint32_t station3;
    uint32_t vectorCount1;      /* #directions / #baselines / #points in the cluster */
    uint32_t vectorCount2;      /* #covariance blocks (G/X/Y), #non-ignored directions (D) */
    uint32_t clusterID;
    uint32_t fileOrder;
    uint32_t sourceFileIndex;
    double term1;               /* measurement value (X|Y|Z component for GNSS) */
    double term2;               /* variance, or XX | XY | XZ */
    double term3;               /* instrument height, or YY | YZ */
    double term4;               /* target height, or ZZ */
    double scale1;              /* phi scalar / derived angle */
    double scale2;              /* lambda scalar / derived-angle variance */
    double scale3;              /* height scalar / derived-angle covariance */
    double scale4;              /* whole-matrix scalar */
    double measAdj;
    double measCorr;
    double measAdjPrec;
    double residualPrec;
    double NStat;
    double TStat;
    double PelzerRel;
    double preAdjCorr;
    double preAdjMeas;
} dna_msr_t;

typedef struct dna_stn_t {
    char stationName[DNA_STN_NAME_WIDTH];
    char stationNameOrig[DNA_STN_NAME_ORIG_WIDTH];
    char stationConst[DNA_STN_CONST_WIDTH];   /* "CCC", "FFF", "CCF", ... (lat, lon, height) */
    char stationType[DNA_STN_TYPE_WIDTH];     /* "LLH", "UTM", "XYZ" */
    uint16_t suppliedStationType;
    double initialLatitude;
    double currentLatitude;                   /* radians */
    double initialLongitude;
    double currentLongitude;                  /* radians */
    double initialHeight;
    double currentHeight;                     /* ellipsoidal, metres */
    uint16_t suppliedHeightRefFrame;
    float geoidSep;
    float geoidSepUnc;
    double meridianDef;
    double verticalDef;
    int16_t zone;
    char description[DNA_STN_DESC_WIDTH];
    uint32_t fileOrder;
    uint32_t nameOrder;
    uint32_t clusterID;
    uint16_t unusedStation;
    char epsgCode[DNA_STN_EPSG_WIDTH];
    char epoch[DNA_STN_EPOCH_WIDTH];
    char observation_epoch[DNA_STN_EPOCH_WIDTH];
    char plate[DNA_STN_PLATE_WIDTH];
} dna_stn_t;

#ifdef __cplusplus
}
static_assert(sizeof(dna_msr_t) == 208, "msr_t must be 208 bytes");
static_assert(offsetof(dna_msr_t, ignore) == 38, "msr_t.ignore");
static_assert(offsetof(dna_msr_t, station1) == 40, "msr_t.station1");
static_assert(offsetof(dna_msr_t, vectorCount1) == 52, "msr_t.vectorCount1");
static_assert(offsetof(dna_msr_t, clusterID) == 60, "msr_t.clusterID");
static_assert(offsetof(dna_msr_t, term1) == 72, "msr_t.term1");
static_assert(offsetof(dna_msr_t, scale1) == 104, "msr_t.scale1");
static_assert(offsetof(dna_msr_t, measAdj) == 136, "msr_t.measAdj");
static_assert(offsetof(dna_msr_t, preAdjMeas) == 200, "msr_t.preAdjMeas");
static_assert(sizeof(dna_stn_t) == 352, "stn_t must be 352 bytes");
static_assert(offsetof(dna_stn_t, stationConst) == 71, "stn_t.stationConst");
static_assert(offsetof(dna_stn_t, suppliedStationType) == 80, "stn_t.suppliedStationType");
static_assert(offsetof(dna_stn_t, currentLatitude) == 96, "stn_t.currentLatitude");
static_assert(offsetof(dna_stn_t, currentLongitude) == 112, "stn_t.currentLongitude");
static_assert(offsetof(dna_stn_t, currentHeight) == 128, "stn_t.currentHeight");
static_assert(offsetof(dna_stn_t, geoidSep) == 140, "stn_t.geoidSep");
static_assert(offsetof(dna_stn_t, geoidSepUnc) == 144, "stn_t.geoidSepUnc");
static_assert(offsetof(dna_stn_t, meridianDef) == 152, "stn_t.meridianDef");
static_assert(offsetof(dna_stn_t, verticalDef) == 160, "stn_t.verticalDef");
static_assert(offsetof(dna_stn_t, fileOrder) == 300, "stn_t.fileOrder");
static_assert(offsetof(dna_stn_t, epsgCode) == 314, "stn_t.epsgCode");
#endif

#endif /* DNA_RECORDS_H_ */
