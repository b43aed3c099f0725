// radiation.cu — the solar-radiation raster loop (second half):
//   radiation::computeRadiationDEM      agrolib/solarRadiation/solarRadiation.cpp:1045-1069
//     computeRadiationDemPoint          :946-981
//     computeRadiationRsun              :700-834   (isIlluminated :533, computeShadow :543-604, separateTransmissivity_Erbs_Reindl
//                                                   :634-690, clearSkyBeamHorizontal / clearSkyDiffuseHorizontal :343-392 (Rigollier
//                                                   et al. 2000), getBeamInclined :399, getDiffuseInclined_Muneer :468-519,
//                                                   getReflectedIrradiance :522)
//     computeSunPosition                :1092-1135 over RSUN_compute_solar_position (sunPosition.cpp:35-99) = NREL SOLPOS
//                                       (solPos.cpp: dom2doy :373, geometry :426-586, zen_no_ref :588, ssha :620, tst :682, srss :708,
//                                       sazm :732, refrac :767, amass :816, etr :851, tilt :899)
//     pressureFromAltitude              agrolib/mathFunctions/physics.cpp:39-47
// One thread per DEM cell. Everything of SOLPOS that depends only on the time (day angle, earth radius vector, julian day,
// ecliptic coordinates, declination, right ascension, Greenwich mean sidereal time) is computed ONCE on the host with the
// reference's own float / double mix; the kernel does the per-cell part (local sidereal time from the cell's longitude,
// hour angle, zenith, sunset hour angle, azimuth, refraction from the cell's pressure, air mass, incidence on the cell's slope /
// aspect), the shadow ray march through the resident DEM and the irradiance model, operation by operation (-fmad=false).
// The reference calls acos() with FLOAT arguments in four places: with <math.h> in C++ that is acosf(), whose glibc algorithm
// (fdlibm, not correctly rounded) is restated here bit for bit (checked against glibc 2.39 on 3e7 arguments, oracle header);
// the one powf() (optical air mass) is taken as the correctly rounded float of pow() — 6e-4 of the arguments differ from glibc
// by one ulp, far inside the tolerance. Double-precision sin / cos / tan / exp / pow are CUDA's (<= 2 ulp), whose results
// are rounded to float right away by the algorithm.
// Bound: FP64 pipe (about 400 FP64 operations and a dozen transcendental calls per cell) + the DEM march of the shadow test.
#include <math.h>
#include <string.h>
#include <string>

#include "context.cuh"
#include "nvtx.cuh"

namespace pb {

namespace {

constexpr double kDegRad = 57.295779513;      // solPos.cpp:121 (degrad)
constexpr double kRadDeg = 0.0174532925;      // solPos.cpp:122 (raddeg)
constexpr double kDegToRad = 0.01745329252;   // DEG_TO_RAD, commonConstants.h:255
constexpr double kRadToDeg = 57.295779513;    // RAD_TO_DEG
constexpr double kPi = 3.1415926535898;       // PI, commonConstants.h:249

// what geometry() / dom2doy() leave in SolPosData that does not depend on the cell
struct RadTime {
    int hour, minute, second;
    float timezone;
    float erv, declin, rascen, gmst;
    float cd, sd;                  // localtrig: float(cos / sin (raddeg * declin))
    float localSeconds;            // float(localTime.time): isIlluminated
};

struct RadParams {
    int realSky, realSkyAlgorithm, shadowing, tiltFixed;
    float fixedTilt, fixedAspect;
    float linke, albedo, clearSky;
    float demMaximum;              // myDem.maximum (gis::updateMinMaxRasterGrid of the DEM): bound of the shadow march
    float outFlag;
};

// glibc 2.39 acosf (sysdeps/ieee754/flt-32/e_acosf.c, the fdlibm single-precision port): float arithmetic, IEEE sqrt and
// division, no contraction
__device__ float acosfGlibc(float x)
{
    const float one = 1.0000000000e+00f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
                pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
                pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
                qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    const int hx = __float_as_int(x), ix = hx & 0x7fffffff;
    float z, p, q, r, w, s, c, df;
    if (ix >= 0x3f800000) {
        if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
        return (x - x) / (x - x);
    }
    if (ix < 0x3f000000) {
        if (ix <= 0x23000000) return pio2_hi + pio2_lo;
        z = x * x;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (hx < 0) {
        z = (one + x) * 0.5f;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        s = sqrtf(z);
        r = p / q;
        w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    } else {
        z = (one - x) * 0.5f;
        s = sqrtf(z);
        df = __int_as_float(__float_as_int(s) & 0xfffff000);
        c = (z - df * df) / (s + df);
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        w = r * s + c;
        return 2.0f * (df + w);
    }
}

struct SunPosition {               // TsunPosition, radiationDefinitions.h:70-83
    float rise, set, azimuth, elevation, elevationRefr, incidence, relOptAirMass, relOptAirMassCorr, extraIrradianceNormal,
          extraIrradianceHorizontal;
    bool shadow;
};

// the per-cell part of S_solpos (solPos.cpp:187-237) + computeSunPosition's copy-out (solarRadiation.cpp:1117-1134);
// false where validate() (:293-347) rejects the cell's inputs
__device__ bool sunPositionCell(const RadTime& t, float longitude, float latitude, float temp, float press, float aspect, float tilt,
                                SunPosition& sp)
{
    if (fabsf(longitude) > 180.f || fabsf(latitude) > 90.f) return false;
    if (fabs(double(temp)) > 100.0) return false;
    if (double(press) < 0.0 || double(press) > 2000.0) return false;
    if (fabs(double(tilt)) > 180.0 || fabs(double(aspect)) > 360.0) return false;

    // geometry(), last part: local mean sidereal time and hour angle (:560-578)
    float lmst = t.gmst * 15.f + longitude;
    lmst -= float(360.0 * (int)(double(lmst) / 360.0));
    if (lmst < 0.) lmst += 360.0;
    float hrang = lmst - t.rascen;
    if (double(hrang) < -180.0) hrang += 360.0;
    else if (double(hrang) > 180.0) hrang -= 360.0;

    // localtrig (:869-892)
    const float cd = t.cd, sd = t.sd;
    const float ch = float(cos(kRadDeg * hrang));
    const float cl = float(cos(kRadDeg * latitude));
    const float sl = float(sin(kRadDeg * latitude));

    // zen_no_ref (:588-611)
    float cz = sd * sl + cd * cl * ch;
    if (fabsf(cz) > 1.0) cz = (cz >= 0.0) ? 1.0 : -1.0;
    float zenetr = float(acosfGlibc(cz) * kDegRad);
    if (zenetr > 99.0) zenetr = 99.0;
    const float elevetr = 90.f - zenetr;

    // ssha (:620-645)
    float ssha;
    const float cdcl = cd * cl;
    if (fabsf(cdcl) >= 0.001) {
        const float cssha = -sl * sd / cdcl;
        if (cssha < -1.0) ssha = 180.0;
        else if (cssha > 1.0) ssha = 0.0;
        else ssha = float(kDegRad * acosfGlibc(cssha));
    } else if (((t.declin >= 0.0) && (latitude > 0.0)) || ((t.declin < 0.0) && (latitude < 0.0))) ssha = 180.0;
    else ssha = 0.0;

    // tst (:682-702), interval = 0
    const float tst = (180.f + hrang) * 4.f;
    float tstfix = tst - (float)t.hour * 60.f - t.minute - (float)t.second / 60.f + (float)0 / 120.f;
    while (tstfix > 720.0) tstfix -= 1440.0;
    while (tstfix < -720.0) tstfix += 1440.0;

    // srss (:708-723)
    float sretr, ssetr;
    if (ssha <= 1.0) { sretr = 2999.0; ssetr = -2999.0; }
    else if (ssha >= 179.0) { sretr = -2999.0; ssetr = 2999.0; }
    else {
        sretr = float(720.0 - 4.0 * ssha - tstfix);
        ssetr = float(720.0 + 4.0 * ssha - tstfix);
    }

    // sazm (:732-757)
    const float ce = float(cos(kRadDeg * elevetr));
    const float se = float(sin(kRadDeg * elevetr));
    float azim = 180.0;
    const float cecl = ce * cl;
    if (fabsf(cecl) >= 0.001) {
        float ca = (se * sl - sd) / cecl;
        if (ca > 1.0) ca = 1.0;
        else if (ca < -1.0) ca = -1.0;
        azim = 180.f - float(acosfGlibc(ca) * kDegRad);
        if (hrang > 0) azim = 360.f - azim;
    }

    // refrac (:767-806)
    double refcor;
    if (elevetr > 85.0) refcor = 0.0;
    else {
        const double tanelev = tan(kRadDeg * elevetr);
        if (elevetr >= 5.0) refcor = 58.1 / tanelev - 0.07 / (pow(tanelev, 3.0)) + 0.000086 / (pow(tanelev, 5.0));
        else if (elevetr >= -0.575)
            refcor = 1735.0 + elevetr * (-518.2 + elevetr * (103.4 + elevetr * (-12.79 + elevetr * 0.711)));
        else refcor = -20.774 / tanelev;
        const double prestemp = (press * 283.0) / (1013.0 * (273.0 + temp));
        refcor *= float(prestemp / 3600.0);
    }
    float elevref = float(elevetr + refcor);
    if (elevref < -9.0) elevref = -9.0;
    const float zenref = float(90.0 - elevref);
    const float coszen = float(cos(kRadDeg * zenref));

    // amass (:816-830)
    float amass, ampress;
    if (zenref > 93.0) { amass = -1.0; ampress = -1.0; }
    else {
        const float pw = float(pow(double(96.07995f - zenref), double(-1.6364f)));      // powf in the reference (see the header)
        amass = 1.0f / float(cos(kRadDeg * zenref) + 0.50572f * pw);
        ampress = amass * press / 1013.0f;
    }

    // etr (:851-861), solcon = 1367 (S_init :281)
    float etrn, etr;
    if (coszen > 0.0) { etrn = 1367.0f * t.erv; etr = etrn * coszen; }
    else { etrn = 0.0; etr = 0.0; }

    // tilt (:899-926)
    const double ca2 = cos(kRadDeg * azim), cp = cos(kRadDeg * aspect), ct = cos(kRadDeg * tilt);
    const double sa2 = sin(kRadDeg * azim), spn = sin(kRadDeg * aspect), stl = sin(kRadDeg * tilt);
    const double sz = sin(kRadDeg * zenref);
    const float cosinc = float(coszen * ct + sz * stl * (ca2 * cp + sa2 * spn));

    // computeSunPosition (:1117-1134)
    sp.relOptAirMass = amass;
    sp.relOptAirMassCorr = ampress;
    sp.azimuth = azim;
    sp.elevation = elevetr;
    sp.elevationRefr = elevref;
    sp.extraIrradianceHorizontal = etr;
    sp.extraIrradianceNormal = etrn;
    const double inc = kRadToDeg * ((kPi / 2.0) - acosfGlibc(cosinc));
    sp.incidence = float(inc > 0. ? inc : 0.);
    sp.rise = sretr * 60.f;
    sp.set = ssetr * 60.f;
    sp.shadow = false;
    return true;
}

// computeShadow (:543-604): the ray towards the sun through the DEM, in doubles
__device__ bool shadowDev(const DevGrid& dem, double demMaximum, double x0, double y0, double z0, const SunPosition& sp)
{
    const double cellSize = dem.cellSize;
    const double sinAz = sin(sp.azimuth * kDegToRad);
    const double cosAz = cos(sp.azimuth * kDegToRad);
    const double sinElev = sin(sp.elevationRefr * kDegToRad);
    const double cosElev = cos(sp.elevationRefr * kDegToRad);
    const double tgElev = sinElev / (cosElev > 1e-6 ? cosElev : 1e-6);
    const double stepX = 1.0 * sinAz * cellSize;
    const double stepY = 1.0 * cosAz * cellSize;
    const double stepZ = 1.0 * cellSize * tgElev;
    const double maxDeltaH = cellSize * 1.0 * 2.0;
    double maxDistCount;
    if (fabs(stepZ) < 1e-6) maxDistCount = (demMaximum - z0) / kEpsilon;
    else maxDistCount = (demMaximum - z0) / stepZ;
    double stepCount = 0.0, step = 1.0;
    while (stepCount < maxDistCount) {
        stepCount += step;
        const double x = x0 + stepX * stepCount;
        const double y = y0 + stepY * stepCount;
        const double z = z0 + stepZ * stepCount;
        // Crit3DRasterGrid::getRowCol (gis.cpp:485-492), gis::isOutOfGridRowCol (:771-776)
        const int r = int((y - dem.lly) * dem.invCellSize);
        const int row = (dem.nrRows - 1) - r;
        const int col = int((x - dem.llx) * dem.invCellSize);
        if (unsigned(row) >= unsigned(dem.nrRows) || unsigned(col) >= unsigned(dem.nrCols)) return false;
        const double zDEM = double(__ldg(dem.data + (size_t)row * dem.nrCols + col));
        if (zDEM != double(dem.flag)) {
            if ((zDEM - z) > 0.5) return true;
            step = (z - zDEM) / maxDeltaH;
            if (step < 1.0) step = 1.0;
        }
    }
    return false;
}

__device__ __forceinline__ double clampD(double v, double lo, double hi) { return v < lo ? lo : (hi < v ? hi : v); }   // std::clamp

// separateTransmissivity_Erbs_Reindl (:634-690)
__device__ void separateErbsReindl(double clearSkyTransmissivity, double transmissivity, double sunElevationDeg, double& td, double& Tt)
{
    Tt = clampD(transmissivity, 1e-6, clearSkyTransmissivity);
    if (clearSkyTransmissivity <= 1e-6) { td = 0.0; return; }
    double Kt = Tt / clearSkyTransmissivity;
    Kt = clampD(Kt, 0.0, 1.2);
    const double elevRad = sunElevationDeg * kDegToRad;
    const double se = sin(elevRad);
    const double sinElev = se > 1e-4 ? se : 1e-4;
    double Kd;
    if (Kt <= 0.22) Kd = 1.0 - 0.09 * Kt;
    else if (Kt <= 0.80) Kd = 0.9511 - 0.1604 * Kt + 4.388 * Kt * Kt - 16.638 * Kt * Kt * Kt + 12.336 * Kt * Kt * Kt * Kt;
    else Kd = 0.165;
    double KdReindl = Kd;
    if (sunElevationDeg > 0.0) KdReindl = Kd + (0.10 + 0.12 * sunElevationDeg / 90.0) * (1.0 - exp(-1.0 / sinElev));
    KdReindl = clampD(KdReindl, 0.0, 1.0);
    td = Tt * KdReindl;
}

__device__ __forceinline__ unsigned orderedKeyRad(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// computeRadiationDemPoint (:946-981) + computeRadiationRsun (:700-834) for every valid DEM cell
__global__ void __launch_bounds__(128)
radiation_dem_kernel(RadTime t, RadParams P, DevGrid dem, const float* __restrict__ lat, const float* __restrict__ lon,
                     const float* __restrict__ slopeMap, const float* __restrict__ aspectMap, const float* __restrict__ transmissivityMap,
                     float* __restrict__ outElev, float* __restrict__ outBeam, float* __restrict__ outDiffuse,
                     float* __restrict__ outReflected, float* __restrict__ outGlobal, unsigned* __restrict__ minmax /* 5 x {min, max} */)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)dem.nrRows * dem.nrCols;
    float o[5] = {P.outFlag, P.outFlag, P.outFlag, P.outFlag, P.outFlag};
    bool wrote = false;
    if (i < n) {
        const float heightF = dem.data[i];
        if (!isEqualF(heightF, dem.flag)) {
            const int row = int(i / dem.nrCols), col = int(i - (long long)row * dem.nrCols);
            const double height = double(heightF);
            // dem.getXY (gis.cpp:473-477)
            const double x = dem.llx + dem.cellSize * (double(col) + 0.5);
            const double y = dem.lly + dem.cellSize * (double(dem.nrRows - row) - 0.5);
            const double rlat = double(lat[i]), rlon = double(lon[i]);
            const double slope = double(P.tiltFixed ? P.fixedTilt : slopeMap[i]);
            const double aspect = double(P.tiltFixed ? P.fixedAspect : aspectMap[i]);
            double transmissivity = double(transmissivityMap[i]);
            const double linke = double(P.linke), albedo = double(P.albedo), clearSky = double(P.clearSky);
            // computeRadiationRsun
            const double temperature = 10.0;                        // TEMPERATURE_DEFAULT
            // pressureFromAltitude (physics.cpp:39-47) * 0.01
            const double pressure = (101325. * pow(1 + height * 0.0065 / 293.16, -9.80665 / (0.0065 * 287.058))) * 0.01;
            SunPosition sp;
            bool ok = sunPositionCell(t, float(rlon), float(rlat), float(temperature), float(pressure), float(aspect), float(slope), sp);
            double beam = 0, diffuse = 0, reflected = 0, global = 0;
            if (ok) {
                // isIlluminated (:533-541)
                const bool illuminated = (sp.rise != kNoData) && (sp.set != kNoData) && (sp.elevationRefr != kNoData) &&
                                         (t.localSeconds >= sp.rise) && (t.localSeconds <= sp.set) && (sp.elevationRefr > 0);
                if (P.shadowing) {
                    sp.shadow = !illuminated;
                    // dem.isLoaded && !isOutOfGridXY(x, y, header): a cell centre is always inside
                    if (illuminated) sp.shadow = shadowDev(dem, double(P.demMaximum), x, y, height, sp);
                }
                if (illuminated) {
                    if (P.realSky && transmissivity == double(kNoData)) ok = false;
                    else {
                        double Gh, dH, diffuseTransmittance, globalTransmittance;
                        if (P.realSkyAlgorithm == 0) {               // RADIATION_REALSKY_TOTALTRANSMISSIVITY
                            if (!P.realSky) transmissivity = clearSky;
                            separateErbsReindl(clearSky, transmissivity, double(sp.elevationRefr), diffuseTransmittance, globalTransmittance);
                            Gh = sp.extraIrradianceHorizontal * transmissivity;
                            dH = sp.extraIrradianceHorizontal * diffuseTransmittance;
                        } else {
                            // clearSkyBeamHorizontal (:343-359)
                            const double airMass = double(sp.relOptAirMassCorr);
                            double rayleighThickness;
                            if (airMass <= 20)
                                rayleighThickness = 1. / (6.6296 + 1.7513 * airMass - 0.1202 * airMass * airMass + 0.0065 * pow(airMass, 3.0) -
                                                          0.00013 * pow(airMass, 4.0));
                            else rayleighThickness = 1. / (10.4 + 0.718 * airMass);
                            const double Bhc = double(sp.extraIrradianceNormal) * sin(sp.elevationRefr * kDegToRad) *
                                               exp(-0.8662 * linke * airMass * rayleighThickness);
                            // clearSkyDiffuseHorizontal (:366-392)
                            double Dhc = 0;
                            if (!(sp.elevationRefr <= 1e-3)) {
                                double Trd = -0.015843 + linke * (0.030543 + 0.0003797 * linke);
                                Trd = Trd > 1e-6 ? Trd : 1e-6;
                                const double se = sin(sp.elevationRefr * kDegToRad);
                                const double sinElev = se > 1e-5 ? se : 1e-5;
                                double A0 = 0.26463 + linke * (-0.061581 + 0.0031408 * linke);
                                if ((A0 * Trd) < 0.0022) A0 = 0.002 / Trd;
                                const double A1 = 2.0402 + linke * (0.018945 - 0.011161 * linke);
                                const double A2 = -1.3025 + linke * (0.039231 + 0.0085079 * linke);
                                const double Fd = A0 + A1 * sinElev + A2 * sinElev * sinElev;
                                Dhc = sp.extraIrradianceNormal * Fd * Trd;
                            }
                            const double Ghc = Dhc + Bhc;
                            if (P.realSky) {
                                Gh = Ghc * transmissivity / clearSky;
                                separateErbsReindl(clearSky, transmissivity, double(sp.elevationRefr), diffuseTransmittance, globalTransmittance);
                                const double dhsOverGhs = diffuseTransmittance / globalTransmittance;
                                dH = dhsOverGhs * Gh;
                            } else { Gh = Ghc; dH = Dhc; }
                        }
                        double Bh;
                        if (!sp.shadow && sp.incidence > 0.) Bh = Gh - dH;
                        else { Bh = 0; Gh = dH; }
                        if (slope == 0) { beam = Bh; diffuse = dH; reflected = 0; global = Gh; }
                        else {
                            if (!sp.shadow && sp.incidence > 0.) {
                                // getBeamInclined (:399-405)
                                const double se = sin(sp.elevationRefr * kDegToRad);
                                const double sinElev = se > 1e-6 ? se : 1e-6;
                                const double si = sin(sp.incidence * kDegToRad);
                                const double sinIncidence = si > 0.0 ? si : 0.0;
                                beam = Bh * (sinIncidence / sinElev);
                            } else beam = 0;
                            // getDiffuseInclined_Muneer (:468-519)
                            if (sp.elevationRefr < 1e-6) diffuse = 0.0f;
                            else {
                                const double slopeRad = slope * kDegToRad, aspectRad = aspect * kDegToRad;
                                const double elevationRad = sp.elevationRefr * kDegToRad;
                                const double se = sin(elevationRad);
                                const double sinElev = se > 1e-6 ? se : 1e-6;
                                const double sinSlope = sin(slopeRad), cosSlope = cos(slopeRad);
                                double Kb = Bh / (sp.extraIrradianceNormal * sinElev);
                                Kb = clampD(Kb, 0.0, 1.2);
                                const double r_sky = (1.0 + cosSlope) / 2.0;
                                const double Fg = sinSlope - slopeRad * cosSlope - kPi * pow(sin(slope * 0.5 * kDegToRad), 2.0);
                                const bool lowSun = (sp.elevationRefr < 3.0);
                                double Fx;
                                if (sp.shadow || sp.incidence <= 0.1) Fx = r_sky + Fg * 0.252271;
                                else {
                                    const double nn = 0.00263 - Kb * (0.712 + 0.6883 * Kb);
                                    const double termBeam = sin(sp.incidence * kDegToRad) / sinElev;
                                    if (!lowSun) Fx = (nn * Fg + r_sky) * (1.0 - Kb) + Kb * termBeam;
                                    else {
                                        const double azimuthLocalDiff = fmod(sp.azimuth * kDegToRad - aspectRad + 2 * kPi, 2 * kPi);
                                        const double dd = 0.1 - 0.008 * elevationRad;
                                        const double denom2 = dd > 0.05 ? dd : 0.05;
                                        Fx = (nn * Fg + r_sky) * (1.0 - Kb) + Kb * sinSlope * cos(azimuthLocalDiff) / denom2;
                                    }
                                }
                                diffuse = dH * Fx;
                            }
                            // getReflectedIrradiance (:522-530), slope passed as float(radPoint.slope)
                            const double slopeF = double(float(slope));
                            if (slopeF < 1e-6) reflected = 0.;
                            else {
                                const double alb = clampD(albedo, 0.0, 1.0);
                                reflected = (alb * (Bh + dH) * (1. - cos(slopeF * kDegToRad)) / 2.);
                            }
                            global = beam + diffuse + reflected;
                        }
                    }
                }
                // not illuminated: all four are 0 (:753-761)
            }
            if (ok) {
                o[0] = sp.elevationRefr;
                o[1] = float(beam);
                o[2] = float(diffuse);
                o[3] = float(reflected);
                o[4] = float(global);
                wrote = true;
            }
        }
        outElev[i] = o[0];
        outBeam[i] = o[1];
        outDiffuse[i] = o[2];
        outReflected[i] = o[3];
        outGlobal[i] = o[4];
    }
    // gis::updateMinMaxRasterGrid of the five maps (updateRadiationMaps :1072-1079)
#pragma unroll
    for (int k = 0; k < 5; k++) {
        float vmin = INFINITY, vmax = -INFINITY;
        if (wrote && !isEqualF(o[k], P.outFlag) && !isEqualF(o[k], kNoData)) { vmin = o[k]; vmax = o[k]; }
        for (int q = 16; q; q >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, q));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, q));
        }
        if ((threadIdx.x & 31) == 0) {
            if (vmin != INFINITY) atomicMin(&minmax[2 * k], orderedKeyRad(vmin));
            if (vmax != -INFINITY) atomicMax(&minmax[2 * k + 1], orderedKeyRad(vmax));
        }
    }
}

float keyToFloatRad(unsigned k)
{
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// dom2doy (:373-381) + the time-only part of geometry() (:426-559) + localtrig's declination terms, host arithmetic in the
// reference's float / double mix (same libm as the reference's own build)
bool radTimeOf(const pb_radiation_settings& s, RadTime& t)
{
    static const int monthDays[13] = {0, 0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
    const int hour = s.secondsOfDay / 3600, minute = (s.secondsOfDay - hour * 3600) / 60, second = s.secondsOfDay - hour * 3600 - minute * 60;
    // validate (:293-330)
    if (s.year < 1950 || s.year > 2100 || s.month < 1 || s.month > 12 || s.day < 1 || s.day > 31) return false;
    if (hour < 0 || hour > 24 || minute < 0 || minute > 59 || second < 0 || second > 59) return false;
    if ((hour == 24 && minute > 0) || (hour == 24 && second > 0)) return false;
    const float timezone = float(s.timeZone);
    if (fabs(timezone) > 12.f) return false;
    int daynum = s.day + monthDays[s.month];
    if (((s.year % 4) == 0) && (((s.year % 100) != 0) || ((s.year % 400) == 0)) && (s.month > 2)) daynum += 1;
    const double degrad = kDegRad, raddeg = kRadDeg;
    const float dayang = float(360.0 * (daynum - 1) / 365.0);
    const double sd = sin(raddeg * dayang), cd = cos(raddeg * dayang);
    const double d2 = 2.0 * dayang;
    const double c2 = cos(raddeg * d2), s2 = sin(raddeg * d2);
    float erv = float(1.000110 + 0.034221 * cd + 0.001280 * sd);
    erv += float(0.000719 * c2 + 0.000077 * s2);
    float utime = float(hour * 3600.0 + minute * 60.0 + second - 0 / 2.0);
    utime = float(utime / 3600.0 - timezone);
    const double delta = (float)(s.year - 1949);
    const int leap = (int)(delta / 4.0);
    const float julday = float(32916.5 + delta * 365.0 + leap + daynum + utime / 24.0);
    const float ectime = float(julday - 51545.0);
    float mnlong = float(280.460 + 0.9856474 * ectime);
    mnlong -= float(360.0 * (int)(mnlong / 360.0));
    if (mnlong < 0.0) mnlong += 360.0;
    float mnanom = float(357.528 + 0.9856003 * ectime);
    mnanom -= float(360.0 * (int)(mnanom / 360.0));
    if (mnanom < 0.0) mnanom += 360.0;
    float eclong = float(mnlong + 1.915 * sin(mnanom * raddeg) + 0.020 * sin(2.0 * mnanom * raddeg));
    eclong -= float(360.0 * (int)(eclong / 360.0));
    if (eclong < 0.0) eclong += 360.0;
    const float ecobli = float(23.439 - 4.0e-07 * ectime);
    const float declin = float(degrad * asin(sin(ecobli * raddeg) * sin(eclong * raddeg)));
    const double top = cos(raddeg * ecobli) * sin(raddeg * eclong);
    const double bottom = cos(raddeg * eclong);
    float rascen = float(degrad * atan2(top, bottom));
    if (rascen < 0.0) rascen += 360.0;
    float gmst = 6.697375f + 0.0657098242f * ectime + utime;
    gmst -= float(24.0 * (int)(gmst / 24.0));
    if (gmst < 0.0) gmst += 24.0;
    t.hour = hour; t.minute = minute; t.second = second;
    t.timezone = timezone;
    t.erv = erv; t.declin = declin; t.rascen = rascen; t.gmst = gmst;
    t.cd = float(cos(raddeg * declin));
    t.sd = float(sin(raddeg * declin));
    t.localSeconds = float(s.secondsOfDay);
    return true;
}

int failR(pb_context* ctx, int code, const std::string& msg) { return pb::failCtx(ctx, code, msg); }

RasterEntry* entryOfR(pb_context* ctx, pb_raster_id id)
{
    if (id < 0 || id >= (int)ctx->rasters.size() || !ctx->rasters[id].used) return nullptr;
    return &ctx->rasters[id];
}

}  // namespace

DevGrid devGridOfRaster(const RasterEntry& r);      // engine.cu

}  // namespace pb

using namespace pb;

extern "C" int pb_compute_radiation_dem(pb_context* ctx, const pb_radiation_settings* s, const pb_radiation_maps* maps,
                                        pb_radiation_out* out)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!s || !maps || !out) return failR(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    RasterEntry* dem = entryOfR(ctx, maps->dem);
    RasterEntry* lat = entryOfR(ctx, maps->lat);
    RasterEntry* lon = entryOfR(ctx, maps->lon);
    RasterEntry* tr = entryOfR(ctx, maps->transmissivity);
    RasterEntry* slope = entryOfR(ctx, maps->slope);
    RasterEntry* aspect = entryOfR(ctx, maps->aspect);
    const bool tiltFixed = s->tiltMode == 1;
    if (!dem || !lat || !lon || !tr || (!tiltFixed && (!slope || !aspect))) return failR(ctx, PB_ERR_INVALID, "a radiation input map is not a live raster");
    for (RasterEntry* e : {lat, lon, tr, slope, aspect})
        if (e && (e->h.nrRows != dem->h.nrRows || e->h.nrCols != dem->h.nrCols))
            return failR(ctx, PB_ERR_INVALID, "the radiation maps must have the DEM's geometry (they are read by row / col)");
    RadTime t;
    if (!radTimeOf(*s, t)) return failR(ctx, PB_ERR_INVALID, "date / time outside what the solar position algorithm accepts (solPos.cpp:293-330)");
    RadParams P{};
    P.realSky = s->realSky; P.realSkyAlgorithm = s->realSkyAlgorithm; P.shadowing = s->shadowing; P.tiltFixed = tiltFixed;
    P.fixedTilt = s->tilt; P.fixedAspect = s->aspect;
    P.linke = s->linke; P.albedo = s->albedo; P.clearSky = s->clearSky;
    P.demMaximum = s->demMaximum;
    P.outFlag = dem->h.flag;
    const size_t n = dem->n;
    // five output maps: new resident rasters (ids returned) from the pool
    pb_raster_id ids[5];
    float* dOut[5];
    for (int k = 0; k < 5; k++) {
        int slot = -1;
        for (size_t i = 0; i < ctx->rasters.size(); i++) if (!ctx->rasters[i].used) { slot = (int)i; break; }
        if (slot < 0) { ctx->rasters.emplace_back(); slot = (int)ctx->rasters.size() - 1; }
        void* d = nullptr;
        CUDA_TRY(ctx, ctx->pool.acquire(n * sizeof(float), &d));
        dem = entryOfR(ctx, maps->dem);             // the vector may have grown
        RasterEntry& e = ctx->rasters[slot];
        e = RasterEntry();
        e.d = static_cast<float*>(d);
        e.h = dem->h;
        e.n = n;
        e.used = true;
        if (ctx->outInitRaster == slot) ctx->outInitRaster = -1;
        ids[k] = slot;
        dOut[k] = e.d;
    }
    dem = entryOfR(ctx, maps->dem); lat = entryOfR(ctx, maps->lat); lon = entryOfR(ctx, maps->lon); tr = entryOfR(ctx, maps->transmissivity);
    slope = entryOfR(ctx, maps->slope); aspect = entryOfR(ctx, maps->aspect);
    unsigned hmm[10];
    for (int k = 0; k < 5; k++) { hmm[2 * k] = 0xffffffffu; hmm[2 * k + 1] = 0u; }
    CUDA_TRY(ctx, ctx->dScratch.reserve(256));
    unsigned* dmm = reinterpret_cast<unsigned*>(ctx->dScratch.p);
    CUDA_TRY(ctx, cudaMemcpyAsync(dmm, hmm, sizeof(hmm), cudaMemcpyHostToDevice, ctx->stream));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    radiation_dem_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(t, P, devGridOfRaster(*dem), lat->d, lon->d, slope ? slope->d : nullptr,
                                                                              aspect ? aspect->d : nullptr, tr->d, dOut[0], dOut[1], dOut[2],
                                                                              dOut[3], dOut[4], dmm);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(hmm, dmm, sizeof(hmm), cudaMemcpyDeviceToHost, ctx->stream));
    pb_raster_out* outs[5] = {out->sunElevation, out->beam, out->diffuse, out->reflected, out->global};
    const int rows = dem->h.nrRows, cols = dem->h.nrCols;
    for (int k = 0; k < 5; k++) {
        pb_raster_out* o = outs[k];
        if (!o || (!o->data && !o->rows)) continue;
        if (o->data) CUDA_TRY(ctx, cudaMemcpyAsync(o->data, dOut[k], n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        else
            for (int r = 0; r < rows; r++)
                CUDA_TRY(ctx, cudaMemcpyAsync(o->rows[r], dOut[k] + size_t(r) * cols, size_t(cols) * sizeof(float), cudaMemcpyDeviceToHost,
                                              ctx->stream));
        ctx->timing.d2h_bytes += (int64_t)(n * sizeof(float));
    }
    cudaEventRecord(ctx->ev[3], ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[1], ctx->ev[3]);
    for (int k = 0; k < 5; k++) {
        pb_raster_out* o = outs[k];
        if (o) {
            o->nValid = (int64_t)n;
            if (hmm[2 * k] == 0xffffffffu) { o->minimum = kNoData; o->maximum = kNoData; }
            else { o->minimum = keyToFloatRad(hmm[2 * k]); o->maximum = keyToFloatRad(hmm[2 * k + 1]); }
        }
        if (out->keepResident) out->ids[k] = ids[k];
        else {
            out->ids[k] = -1;
            ctx->pool.release(ctx->rasters[ids[k]].d);
            ctx->rasters[ids[k]] = RasterEntry();
        }
    }
    return PB_OK;
}
