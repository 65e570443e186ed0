// jbf_main.cpp -- the JBF leg of the reference's main() (main.cpp:43-67, 160-183) against the drop-in
// headers: allocate device buffers, upload one RGB-D frame, Buffer2D::updateData, JBF.Process,
// read the filtered depth back.  Pure C++ host code over the C ABI (no Python, no torch).
//
//   g++ -std=c++11 -DKDME_NO_OPENCV -Iinclude -I/usr/local/cuda/include examples/jbf_main.cpp \
//       -Lkinectdepthmapenhancement_b200 -lkdme_b200 -L/usr/local/cuda/lib64 -lcudart \
//       -Wl,-rpath,$PWD/kinectdepthmapenhancement_b200 -o jbf_main
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Buffer2D.h"
#include "JointBilateralFilter.h"

int main(int argc, char** argv) {
    const int W = 640, H = 480;                       // Kinect::Width / Kinect::Height (Kinect/Kinect.cpp:10-11)
    const int radius = argc > 1 ? atoi(argv[1]) : 2;  // reference default window 5
    std::vector<float> depth_h(W * H);
    std::vector<unsigned char> color_h(W * H * 3);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const bool left = x < W / 2;
            depth_h[y * W + x] = ((x * 7 + y * 13) % 97 == 0) ? 0.0f : (left ? 1200.0f : 2600.0f) + 0.3f * x + (float)((x * 31 + y * 17) % 11);
            color_h[(y * W + x) * 3 + 0] = left ? 60 : 180;
            color_h[(y * W + x) * 3 + 1] = (unsigned char)(100 + (x + y) % 7);
            color_h[(y * W + x) * 3 + 2] = left ? 200 : 40;
        }
    float *inputDepth_Device = 0, *bufferDepth_Device = 0;
    unsigned char* color_dev = 0;
    if (cudaMalloc(&inputDepth_Device, sizeof(float) * W * H) != cudaSuccess) { std::printf("no CUDA device\n"); return 0; }
    cudaMalloc(&bufferDepth_Device, sizeof(float) * W * H);
    cudaMalloc(&color_dev, W * H * 3);
    cudaMemcpy(inputDepth_Device, depth_h.data(), sizeof(float) * W * H, cudaMemcpyHostToDevice);   // main.cpp:160
    cudaMemcpy(color_dev, color_h.data(), W * H * 3, cudaMemcpyHostToDevice);                       // main.cpp:163 (upload)
    cv::gpu::GpuMat Color_Device(H, W, color_dev, (size_t)W * 3);                                    // createContinuous, main.cpp:62
    try {
        Buffer2D Buffer(W, H);                                                                       // main.cpp:65
        JointBilateralFilter JBF(W, H, 70.0f, 50.0f, 20.0f, radius);                                 // main.cpp:67
        Buffer.updateData(inputDepth_Device);                                                        // main.cpp:99
        Buffer.getDepthMap(bufferDepth_Device);                                                      // main.cpp:104
        JBF.Process(inputDepth_Device, Color_Device);                                                // main.cpp:179
        const float* filtered = JBF.getFiltered_Host();
        int holes_in = 0, holes_out = 0;
        double sum = 0;
        for (int i = 0; i < W * H; i++) { holes_in += depth_h[i] <= 50.0f; holes_out += filtered[i] <= 0.0f; sum += filtered[i]; }
        std::printf("JBF r=%d: holes %d -> %d, mean filtered depth %.3f mm\n", radius, holes_in, holes_out, sum / (W * H));
        if (holes_out != 0) return 2;   // isolated holes must be filled
    } catch (const std::exception& e) {
        std::printf("error: %s\n", e.what());
        return 1;
    }
    cudaFree(inputDepth_Device); cudaFree(bufferDepth_Device); cudaFree(color_dev);
    return 0;
}
