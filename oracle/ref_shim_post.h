/*
 * ref_shim_post.h -- extern "C" drivers placed AFTER the reference's kernel
 * text (see ref_shim_pre.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Each driver replaces a <<<grid, block>>> launch by a loop that calls the
 * kernel body once per pixel with blockDim = (1,1,1), blockIdx = (x,y).  Unlike
 * the reference's launch (grid = W/32 x H/24, JointBilateralFilter.cu:289) the
 * loop covers every pixel.
 */
#undef powf
#undef pow

extern "C" {

#define REF_API __attribute__((visibility("default")))

/* JointBilateralFilter.cu:289-290 */
REF_API void ref_jbf(int width, int height, float *depth, unsigned char *guide, float *spatial,
                     float *out, int window_size, float color_sigma, float depth_sigma, int n_threads)
{
    cv::gpu::GpuMat g; g.data = guide;
#pragma omp parallel for schedule(dynamic, 2) num_threads(n_threads > 0 ? n_threads : 1)
    for (int y = 0; y < height; y++) {
        blockDim = {1, 1, 1}; threadIdx = {0, 0, 0};
        for (int x = 0; x < width; x++) {
            blockIdx = {x, y, 0};
            joint_bilateral_filtering(width, height, depth, g, spatial, out, window_size, color_sigma,
                                      depth_sigma);
        }
    }
}

/* EdgeRefinedSuperpixel.cu:220-221, executed with race-free semantics: every
 * pixel sees the unmodified input (each worker keeps a private copy of the
 * depth plane and restores the one element the kernel overwrote). */
REF_API void ref_guided_fill(int width, int height, const float *depth, unsigned char *guide,
                             int *labels, float *spatial, float *out, int window_size,
                             float color_sigma, float depth_sigma, int n_threads)
{
    cv::gpu::GpuMat g; g.data = guide;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        std::vector<float> priv(depth, depth + (size_t)width * height);
#pragma omp for schedule(dynamic, 2)
        for (int y = 0; y < height; y++) {
            blockDim = {1, 1, 1}; threadIdx = {0, 0, 0};
            for (int x = 0; x < width; x++) {
                blockIdx = {x, y, 0};
                size_t k = (size_t)y * width + x;
                float saved = priv[k];
                depthmap_enhancement(width, height, priv.data(), g, labels, spatial, window_size,
                                     color_sigma, depth_sigma);
                out[k] = priv[k];
                priv[k] = saved;
            }
        }
    }
}

/* MarkovRandomField.cu:45-46 */
REF_API void ref_mrf(int width, int height, float *depth, unsigned char *guide, float *out,
                     int window_size, float color_sigma, float smooth_sigma, int n_threads)
{
    cv::gpu::GpuMat g; g.data = guide;
#pragma omp parallel for schedule(dynamic, 2) num_threads(n_threads > 0 ? n_threads : 1)
    for (int y = 0; y < height; y++) {
        blockDim = {1, 1, 1}; threadIdx = {0, 0, 0};
        for (int x = 0; x < width; x++) {
            blockIdx = {x, y, 0};
            markov_random_field(width, height, depth, g, out, window_size, color_sigma, smooth_sigma);
        }
    }
}

/* Projection_GPU.cu:264-265, race-free: each worker filters a private copy of the cloud and
 * restores the element the kernel overwrote. */
REF_API void ref_depth_bilateral_xyz(float *normalized, const float *in, float *out, float *spatial,
                                     int window_size, float depth_sigma, int width, int height, int n_threads)
{
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        std::vector<float3> priv((const float3 *)in, (const float3 *)in + (size_t)width * height);
#pragma omp for schedule(dynamic, 2)
        for (int y = 0; y < height; y++) {
            blockDim = {1, 1, 1}; threadIdx = {0, 0, 0};
            for (int x = 0; x < width; x++) {
                blockIdx = {x, y, 0};
                size_t k = (size_t)y * width + x;
                float3 saved = priv[k];
                bilateralfilter((float3 *)normalized, priv.data(), spatial, window_size, depth_sigma, width, height);
                ((float3 *)out)[k] = priv[k];
                priv[k] = saved;
            }
        }
    }
}

/* ArrayBuffer.cu:29, Buffer2D.cu:53-56,73-77,91-94,116-120,144-147 */
#define REF_FOR_PIXELS                                   \
    blockDim = {1, 1, 1}; threadIdx = {0, 0, 0};         \
    for (int y = 0; y < height; y++)                     \
        for (int x = 0; x < width; x++) {                \
            blockIdx = {x, y, 0};
#define REF_END }

REF_API void ref_buf_init(float *buf, int width, int height)
{
    REF_FOR_PIXELS initDeviceMemoryElementsKernel((ArrayBuffer::weighted_d *)buf, width, height); REF_END
}
REF_API void ref_buf_insert_f32(float *buf, float *data, int width, int height)
{
    REF_FOR_PIXELS insertDataKernel((ArrayBuffer::weighted_d *)buf, data, width, height, 0, 100.0f); REF_END
}
REF_API void ref_buf_insert_f32x2(float *buf, float *data, int width, int height)
{
    REF_FOR_PIXELS insertDataKernel((ArrayBuffer::weighted_d *)buf, (float2 *)data, width, height, 0); REF_END
}
REF_API void ref_buf_get_depth(float *buf, float *out, int width, int height)
{
    REF_FOR_PIXELS getDepthMapKernel((ArrayBuffer::weighted_d *)buf, out, width, height, 0); REF_END
}
REF_API void ref_buf_get_weight(float *buf, float *out, int width, int height)
{
    REF_FOR_PIXELS getWeightMapKernel((ArrayBuffer::weighted_d *)buf, out, width, height); REF_END
}
REF_API void ref_buf_update(float *buf, float *data, int width, int height)
{
    REF_FOR_PIXELS updateDataKernel((ArrayBuffer::weighted_d *)buf, data, width, height, 0, 100.0f); REF_END
}

} /* extern "C" */
