#ifndef OPENMM_STUB_CUDAARRAY_H_
#define OPENMM_STUB_CUDAARRAY_H_
// Stand-in for OpenMM's CudaArray: a typed device allocation with upload / download.
#include "openmm/OpenMMException.h"
#include <cuda_runtime_api.h>
#include <string>
#include <vector>

typedef unsigned long long CUdeviceptr_stub;

namespace OpenMM {

class CudaContext;

class CudaArray {
public:
    CudaArray() : pointer(nullptr), size(0), elementSize(0) {}
    CudaArray(CudaContext& context, size_t size, int elementSize, const std::string& name) : pointer(nullptr), size(0), elementSize(0) {
        initialize(context, size, elementSize, name);
    }
    ~CudaArray() { if (pointer) cudaFree(pointer); }
    template <class T> static CudaArray* create(CudaContext& context, size_t size, const std::string& name) {
        return new CudaArray(context, size, sizeof(T), name);
    }
    void initialize(CudaContext&, size_t n, int elemSize, const std::string& arrayName) {
        if (pointer) cudaFree(pointer);
        size = n; elementSize = elemSize; name = arrayName;
        if (cudaMalloc(&pointer, n * elemSize > 0 ? n * elemSize : 1) != cudaSuccess) throw OpenMMException("Error creating array " + name);
        cudaMemset(pointer, 0, n * elemSize);
    }
    template <class T> void initialize(CudaContext& context, size_t n, const std::string& arrayName) { initialize(context, n, sizeof(T), arrayName); }
    bool isInitialized() const { return pointer != nullptr; }
    size_t getSize() const { return size; }
    int getElementSize() const { return elementSize; }
    const std::string& getName() const { return name; }
    CUdeviceptr_stub& getDevicePointer() { return reinterpret_cast<CUdeviceptr_stub&>(pointer); }
    template <class T> void upload(const std::vector<T>& data) { upload(data.data()); }
    void upload(const void* data) {
        if (cudaMemcpy(pointer, data, size * elementSize, cudaMemcpyHostToDevice) != cudaSuccess) throw OpenMMException("Error uploading array " + name);
    }
    template <class T> void download(std::vector<T>& data) { data.resize(size); download(data.data()); }
    void download(void* data) {
        if (cudaMemcpy(data, pointer, size * elementSize, cudaMemcpyDeviceToHost) != cudaSuccess) throw OpenMMException("Error downloading array " + name);
    }
private:
    CudaArray(const CudaArray&);
    CudaArray& operator=(const CudaArray&);
    void* pointer;
    size_t size;
    int elementSize;
    std::string name;
};

} // namespace OpenMM
#endif
