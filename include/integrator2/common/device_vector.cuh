// deviceVector<T>: raw device array + size + capacity, same public members and semantics as the reference's
// (/root/reference/src/common/device_vector.cuh:16-109): resize() beyond capacity reallocates 1.25x WITHOUT
// preserving contents; not copyable.
#ifndef DEVICE_VECTOR_CUH
#define DEVICE_VECTOR_CUH

#include "constants.h"
#include "cuda_memory.cuh"

template <class T>
struct deviceVector {
    T *data = nullptr;
    int size = 0;
    int capacity = 0;

    deviceVector() = default;
    deviceVector(const deviceVector &) = delete;
    deviceVector &operator=(const deviceVector &) = delete;
    ~deviceVector() { release(); }

    void allocate(int newSize) {
        release();
        allocate_device(&data, newSize);
        size = capacity = newSize;
    }
    size_t bytes() const { return size * sizeof(T); }
    void swap(deviceVector<T> &other) {
        T *d = other.data; other.data = data; data = d;
        int s = other.size; other.size = size; size = s;
        int c = other.capacity; other.capacity = capacity; capacity = c;
    }
    void resize(int newSize) {
        if (newSize <= capacity) { size = newSize; return; }
        allocate((int)(CONSTANTS::MEMORY_REALLOCATION_COEFFICIENT * newSize));
        size = newSize;
        printf("Reallocation performed\n");
    }

private:
    void release() {
        if (data) { free_device(data); data = nullptr; size = capacity = 0; }
    }
};

#endif  // DEVICE_VECTOR_CUH
