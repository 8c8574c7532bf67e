/* gdal_proj_stubs.c -- link-only definitions of the reference's GDAL / PROJ wrappers (test infrastructure).
 *
 * The reference binds 13 C symbols (src/GeoTiffRead.f90:44-106 from RasterData.cpp, src/utm.f90:55-103 from
 * cUTM.cpp).  With `Topog Type = Function` none of them is called (src/TopogSettings.f90:378-388,
 * src/dem.f90:405-413): RunParams%Georeference stays false, no raster is opened, no coordinate is projected.
 * These definitions let the Fortran sources link without GDAL and PROJ; each aborts if it is ever reached, so a
 * run that would need the real library fails loudly instead of computing with zeros. */
#include <stdio.h>
#include <stdlib.h>

static void unreachable(const char *name) {
   fprintf(stderr, "oracle/ref_build stub reached: %s (this build has no GDAL / PROJ; use Topog Type = Function)\n", name);
   abort();
}

/* src/GeoTiffRead.f90:44-106 */
void MallocDouble(void **ptr, const long *n) { *ptr = calloc((size_t)(*n > 0 ? *n : 1), sizeof(double)); }
void FreeDouble(void **ptr) { free(*ptr); *ptr = NULL; }
void GeoTiffInfo(const char *name, void *raster) { (void)name; (void)raster; unreachable("GeoTiffInfo"); }
void GeoTiffArraySectionRead(const char *path, const char *name, void *raster, int *xoff, int *yoff, int *xsize, int *ysize) {
   (void)path; (void)name; (void)raster; (void)xoff; (void)yoff; (void)xsize; (void)ysize;
   unreachable("GeoTiffArraySectionRead");
}
void BuildDEMVRT_raster(const char *path, const char *srtm, const char *name, const int *epsg, const _Bool *embed, const double *minE,
                        const double *maxE, const double *minN, const double *maxN, const double *xres, const double *yres) {
   (void)path; (void)srtm; (void)name; (void)epsg; (void)embed; (void)minE; (void)maxE; (void)minN; (void)maxN; (void)xres; (void)yres;
   unreachable("BuildDEMVRT_raster");
}
void BuildDEMVRT_srtm(const char *path, const char *srtm, const int *epsg, const double *minE, const double *maxE, const double *minN,
                      const double *maxN, const double *xres, const double *yres) {
   (void)path; (void)srtm; (void)epsg; (void)minE; (void)maxE; (void)minN; (void)maxN; (void)xres; (void)yres;
   unreachable("BuildDEMVRT_srtm");
}

/* src/utm.f90:55-103 */
void *proj_transformer__new(int utm_code) { (void)utm_code; return NULL; }   /* constructed unconditionally, never used */
void proj_transformer__delete(void *self) { (void)self; }
void *proj_transformer__wgs84_to_utm(void *self, double lat, double lon) { (void)self; (void)lat; (void)lon; unreachable("wgs84_to_utm"); return NULL; }
void *proj_transformer__utm_to_wgs84(void *self, double e, double n) { (void)self; (void)e; (void)n; unreachable("utm_to_wgs84"); return NULL; }
int latlon_to_zone_number(double lat, double lon) { (void)lat; return (int)((lon + 180.0) / 6.0) + 1; }
int zone_number_to_central_longitude(int zone) { return (zone - 1) * 6 - 180 + 3; }
int latlon_to_utm_epsg(double lat, double lon) { return (lat >= 0.0 ? 32600 : 32700) + latlon_to_zone_number(lat, lon); }
