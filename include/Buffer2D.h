/*
 * Buffer2D.h -- header-only C++ drop-in for ArrayBuffer / Buffer2D
 * (ArrayBuffer/ArrayBuffer.h:9-45, ArrayBuffer/Buffer2D.h:9-35), forwarding to kdme_b200.h.
 * Caller pattern kept: Buffer2D Buffer(W, H); Buffer.updateData(dev); Buffer.getDepthMap(dev)
 * (main.cpp:65,99,104).  The OpenNI overload insertData(xn::DepthMetaData*) (Buffer2D.cpp:18-32)
 * becomes insertData(const uint16_t* host_depth), the same data without the OpenNI type.
 */
#ifndef KDME_BUFFER2D_H
#define KDME_BUFFER2D_H

#include <stdexcept>
#include <string>

#include "kdme_b200.h"

/* float2: the CUDA vector type when <vector_types.h> is reachable (any CUDA toolkit include path), else an
 * identical plain struct, so that insertData(float2*) -- Buffer2D.h:24 -- keeps its reference spelling. */
#if defined(__CUDACC__) || defined(__VECTOR_TYPES_H__)
#define KDME_FLOAT2 float2
#elif defined(__has_include)
#if __has_include(<vector_types.h>)
#include <vector_types.h>
#define KDME_FLOAT2 float2
#endif
#endif
#ifndef KDME_FLOAT2
struct float2 { float x, y; };
#define KDME_FLOAT2 float2
#endif

class ArrayBuffer {
public:
    struct weighted_d { float d; float w; };   /* ArrayBuffer.h:12-15 */
    virtual ~ArrayBuffer() {}
    virtual void insertData(float* data) = 0;
    virtual void insertData(weighted_d* data) = 0;
    virtual void getDepthMap(float* out) = 0;
    virtual void getWeightMap(float* out) = 0;
    virtual void updateData(float* data) = 0;
    weighted_d* getRawPointer() { return reinterpret_cast<weighted_d*>(buf2d_raw(b_)); }   /* ArrayBuffer.cpp:19-21 */
protected:
    ArrayBuffer(int w, int h, int device, void* stream) : width(w), height(h), b_(0) {
        check(buf2d_create(&b_, w, h, device, stream));
    }
    static void check(int rc) {
        if (rc != KDME_OK) throw std::runtime_error(std::string("kdme_b200: ") + kdme_last_error());
    }
    int width, height;
    buf2d_handle* b_;
};

class Buffer2D : public ArrayBuffer {
public:
    explicit Buffer2D(int width, int height, int device = 0, void* stream = 0) : ArrayBuffer(width, height, device, stream) {}
    ~Buffer2D() { buf2d_destroy(b_); }
    virtual void insertData(float* data) { check(buf2d_insert_f32(b_, data)); }                       /* Buffer2D.cu:53-56 */
    virtual void insertData(weighted_d* data) { check(buf2d_insert_dw(b_, reinterpret_cast<float*>(data))); } /* Buffer2D.cpp:13-15 */
    virtual void getDepthMap(float* out) { check(buf2d_get_depth(b_, out)); }                         /* Buffer2D.cu:73-77 */
    virtual void getWeightMap(float* out) { check(buf2d_get_weight(b_, out)); }                       /* Buffer2D.cu:91-94 */
    virtual void updateData(float* data) { check(buf2d_update_f32(b_, data)); }                       /* Buffer2D.cu:116-120 */
    void updateData(float* data, int n_frames) { check(buf2d_update_batch_f32(b_, data, n_frames)); }
    void insertData(KDME_FLOAT2* data) { check(buf2d_insert_f32x2(b_, reinterpret_cast<float*>(data))); }     /* Buffer2D.h:24, Buffer2D.cu:123-147 */
    template <class Float2> void insertData2(Float2* data) { check(buf2d_insert_f32x2(b_, reinterpret_cast<float*>(data))); } /* same, any {x,y} pair type */
    void insertData(const uint16_t* host_depth) { check(buf2d_update_u16_host(b_, host_depth)); }     /* Buffer2D.cpp:18-32 */
private:
    Buffer2D(const Buffer2D&);
    Buffer2D& operator=(const Buffer2D&);
};

#endif
