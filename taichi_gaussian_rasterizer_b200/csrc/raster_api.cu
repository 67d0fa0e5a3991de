// raster_api.cu — gs_raster_fwd / gs_raster_bwd: validation and kernel-family selection.
#include "raster.cuh"

using namespace gs;

static int check_raster(const GsRasterParams* p, const char* who) {
  if (!p) { set_error("%s: null params", who); return GS_ERR_INVALID; }
  if (p->dtype != GS_F32 && p->dtype != GS_F64) { set_error("%s: dtype must be GS_F32 or GS_F64", who); return GS_ERR_INVALID; }
  if (p->image_width <= 0 || p->image_height <= 0 || p->num_features <= 0 || p->num_points < 0 || p->num_overlaps < 0) {
    set_error("%s: bad sizes (image %dx%d, F=%d, V=%lld, K=%lld)", who, p->image_width, p->image_height,
              p->num_features, (long long)p->num_points, (long long)p->num_overlaps);
    return GS_ERR_INVALID;
  }
  if (p->tile_size != 8 && p->tile_size != 16 && p->tile_size != 32) {
    set_error("%s: tile_size must be 8, 16 or 32, got %d", who, p->tile_size);
    return GS_ERR_UNSUPPORTED;
  }
  const int sx = p->pixel_stride_x, sy = p->pixel_stride_y;
  if (sx <= 0 || sy <= 0 || (p->tile_size * p->tile_size) / (sx * sy) < 32) {  // rasterizer/backward.py:33-34
    set_error("%s: pixel_stride (%d, %d) and tile_size %d must allow at least one warp sized (32) tile", who, sx, sy,
              p->tile_size);
    return GS_ERR_INVALID;
  }
  return GS_OK;
}

extern "C" {

size_t gs_raster_workspace_bytes(const GsRasterParams* p) {
  if (!p) return 0;
  return raster_fast_supported(*p) ? raster_fast_workspace_bytes(*p) : 0;
}

int gs_raster_fwd(const GsRasterParams* p, const void* gaussians2d, const void* features, const int32_t* tile_ranges,
                  const int32_t* overlap_to_point, void* image, void* image_alpha, void* visibility, void* workspace,
                  size_t workspace_bytes, void* stream) {
  int rc = check_raster(p, "gs_raster_fwd");
  if (rc != GS_OK) return rc;
  GS_CHECK_ARG(tile_ranges && image && image_alpha, "gs_raster_fwd: null tensor");
  GS_CHECK_ARG(p->num_points == 0 || (gaussians2d && features), "gs_raster_fwd: null gaussians / features");
  GS_CHECK_ARG(p->num_overlaps == 0 || overlap_to_point, "gs_raster_fwd: null overlap_to_point");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = p->dtype == GS_F32 ? 4 : 8;
  if (p->compute_visibility && visibility && p->num_points > 0)
    GS_CUDA(cudaMemsetAsync(visibility, 0, (size_t)p->num_points * es, st));
  RasterArgs a = {};
  a.gaussians2d = gaussians2d; a.features = features; a.tile_ranges = tile_ranges;
  a.overlap_to_point = overlap_to_point; a.image = image; a.image_alpha = image_alpha; a.visibility = visibility;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  if (raster_fast_supported(*p)) {
    const size_t need = raster_fast_workspace_bytes(*p);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("gs_raster_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
      return GS_ERR_WORKSPACE;
    }
    return raster_fwd_fast(*p, a, st);
  }
  return raster_fwd_generic(*p, a, st);
}

int gs_raster_bwd(const GsRasterParams* p, const void* gaussians2d, const void* features, const int32_t* tile_ranges,
                  const int32_t* overlap_to_point, const void* image, const void* grad_image, void* grad_gaussians,
                  void* grad_features, void* point_heuristic, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_raster(p, "gs_raster_bwd");
  if (rc != GS_OK) return rc;
  GS_CHECK_ARG(tile_ranges && grad_image && (image || !p->use_alpha_blending), "gs_raster_bwd: null tensor");
  GS_CHECK_ARG(p->num_points == 0 || (gaussians2d && features), "gs_raster_bwd: null gaussians / features");
  GS_CHECK_ARG(p->num_overlaps == 0 || overlap_to_point, "gs_raster_bwd: null overlap_to_point");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = p->dtype == GS_F32 ? 4 : 8;
  if (grad_gaussians && p->num_points > 0) GS_CUDA(cudaMemsetAsync(grad_gaussians, 0, (size_t)p->num_points * 7 * es, st));
  if (grad_features && p->num_points > 0)
    GS_CUDA(cudaMemsetAsync(grad_features, 0, (size_t)p->num_points * p->num_features * es, st));
  RasterArgs a = {};
  a.gaussians2d = gaussians2d; a.features = features; a.tile_ranges = tile_ranges;
  a.overlap_to_point = overlap_to_point; a.image_in = image; a.grad_image = grad_image;
  a.grad_gaussians = grad_gaussians; a.grad_features = grad_features; a.point_heuristic = point_heuristic;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  if (raster_fast_supported(*p)) {
    const size_t need = raster_fast_workspace_bytes(*p);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("gs_raster_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
      return GS_ERR_WORKSPACE;
    }
    return raster_bwd_fast(*p, a, st);
  }
  return raster_bwd_generic(*p, a, st);
}

}  // extern "C"
