// networks/GridNetwork.hpp -- drop-in for the reference's SPRL::GridNetwork
// (/root/reference/cpp/src/networks/GridNetwork.hpp:26-159): same template signature and
// constructor (a TorchScript path, or "random" to load nothing).  The reference moves the
// module to torch::kCPU and fills a host tensor element by element (:70-97); here the module
// lives on the engine's GPU and is run directly on the leaf-batch buffer the search kernel
// wrote (torch::from_blob, no copy, no host round trip).  exp / mask / normalise of the
// outputs (:107-142) happen in the next search launch, on the device.
#ifndef SPRL_GRID_NETWORK_HPP
#define SPRL_GRID_NETWORK_HPP

#include "../sprl/veneer.hpp"

#include <torch/script.h>
#include <torch/torch.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

namespace SPRL {

template <int NUM_ROWS, int NUM_COLS, int HISTORY_SIZE, int ACTION_SIZE>
class GridNetwork : public INetwork<GridState<NUM_ROWS * NUM_COLS, HISTORY_SIZE>, ACTION_SIZE> {
public:
    using State = GridState<NUM_ROWS * NUM_COLS, HISTORY_SIZE>;

    GridNetwork(std::string path) : m_path(path) {
        if (path == "random") return;       // requested random network instead, nothing to load
        try {
            m_model = std::make_shared<torch::jit::Module>(torch::jit::load(path));
        } catch (const c10::Error& e) {
            std::cerr << "Error loading the model: " << e.what() << std::endl;
        }
    }

    int evaluatorKind() const override { return SPRL_EVAL_EXTERNAL; }

    int prepare(int device, int64_t batch, int planes, int rows, int cols, int actions,
                float** d_in, float** d_logits, float** d_value) override {
        if (!m_model) return -1;
        m_device = torch::Device(torch::kCUDA, (c10::DeviceIndex)device);
        m_model->to(m_device);
        m_model->eval();
        // the reference's worker evaluates in fp32; keep cuDNN / cuBLAS off TF32
        at::globalContext().setAllowTF32CuDNN(false);
        at::globalContext().setAllowTF32CuBLAS(false);
        auto opts = torch::TensorOptions().dtype(torch::kFloat32).device(m_device);
        m_input = torch::zeros({ batch, planes, rows, cols }, opts);
        m_logits = torch::zeros({ batch, actions }, opts);
        m_value = torch::zeros({ batch }, opts);
        *d_in = m_input.data_ptr<float>();
        *d_logits = m_logits.data_ptr<float>();
        *d_value = m_value.data_ptr<float>();
        return 0;
    }

    int forward(const float* d_in, int64_t batch, float* d_logits, float* d_value, void* stream) override {
        (void)batch;
        if (d_in != m_input.data_ptr<float>() || d_logits != m_logits.data_ptr<float>() || d_value != m_value.data_ptr<float>()) return -2;
        torch::NoGradGuard no_grad;
        c10::cuda::CUDAStream s = c10::cuda::getStreamFromExternal((cudaStream_t)stream, m_device.index());
        c10::cuda::CUDAStreamGuard guard(s);
        auto output = m_model->forward({ m_input }).toTuple();
        m_logits.copy_(output->elements()[0].toTensor());
        m_value.copy_(output->elements()[1].toTensor().reshape({ -1 }));
        return 0;
    }

private:
    std::string m_path;
    torch::Device m_device { torch::kCPU };
    std::shared_ptr<torch::jit::script::Module> m_model;
    torch::Tensor m_input, m_logits, m_value;
};

}  // namespace SPRL
#endif
