// networks/GridNetwork.hpp -- drop-in for the reference's SPRL::GridNetwork
// (/root/reference/cpp/src/networks/GridNetwork.hpp:26-159): same template signature and
// constructor (a TorchScript path, or "random" to load nothing).  The reference moves the
// module to torch::kCPU and fills a host tensor element by element (:70-97); here the module
// lives on the engine's GPU and is run directly on the leaf-batch buffer the search kernel
// wrote (no copy, no host round trip).  For boards up to 8x8 the forward pass itself is the library's
// tcgen05 kernel (sprl_evalnet_*, csrc/evalnet.cu) fed with the module's parameters; other
// boards, or SPRL_EVALUATOR=libtorch, run the TorchScript module through LibTorch/cuDNN.
// exp / mask / normalise of the outputs (:107-142) happen in the next search launch, on the device.
#ifndef SPRL_GRID_NETWORK_HPP
#define SPRL_GRID_NETWORK_HPP

#include "../sprl/veneer.hpp"

#include <torch/script.h>
#include <torch/torch.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cstdlib>
#include <cstring>
#include <map>

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

    // Moves the module to the engine's GPU and picks the forward implementation.
    int attach(int device, int planes, int rows, int cols, int actions) override {
        if (!m_model) return -1;
        m_device = torch::Device(torch::kCUDA, (c10::DeviceIndex)device);
        m_model->to(m_device);
        m_model->eval();
        // the reference's worker evaluates in fp32; keep cuDNN / cuBLAS off TF32
        at::globalContext().setAllowTF32CuDNN(false);
        at::globalContext().setAllowTF32CuBLAS(false);
        m_planes = planes; m_rows = rows; m_cols = cols; m_actions = actions;
        const char* ev = std::getenv("SPRL_EVALUATOR");
        if (!m_evalnet && !(ev && std::strcmp(ev, "libtorch") == 0)) createEvalnet(device, planes, rows, cols, actions);
        return 0;
    }

    int prepare(int device, int64_t batch, int planes, int rows, int cols, int actions,
                float** d_in, float** d_logits, float** d_value) override {
        if (attach(device, planes, rows, cols, actions) != 0) return -1;
        auto opts = torch::TensorOptions().dtype(torch::kFloat32).device(m_device);
        m_input = torch::zeros({ batch, planes, rows, cols }, opts);
        m_logits = torch::zeros({ batch, actions }, opts);
        m_value = torch::zeros({ batch }, opts);
        *d_in = m_input.data_ptr<float>();
        *d_logits = m_logits.data_ptr<float>();
        *d_value = m_value.data_ptr<float>();
        return 0;
    }

    ~GridNetwork() { if (m_evalnet) sprl_evalnet_destroy(m_evalnet); }

    // d_in / d_logits / d_value may be any rows of a leaf-batch buffer on this GPU (a match gives each network
    // its half of one buffer).
    int forward(const float* d_in, int64_t batch, float* d_logits, float* d_value, void* stream) override {
        if (m_evalnet) return sprl_evalnet_forward_counted(m_evalnet, d_in, m_dRows, batch, d_logits, d_value, stream);
        if (!m_model || m_planes == 0) return -2;
        torch::NoGradGuard no_grad;
        c10::cuda::CUDAStream s = c10::cuda::getStreamFromExternal((cudaStream_t)stream, m_device.index());
        c10::cuda::CUDAStreamGuard guard(s);
        auto opts = torch::TensorOptions().dtype(torch::kFloat32).device(m_device);
        torch::Tensor in = torch::from_blob(const_cast<float*>(d_in), { batch, m_planes, m_rows, m_cols }, opts);
        torch::Tensor logits = torch::from_blob(d_logits, { batch, m_actions }, opts);
        torch::Tensor value = torch::from_blob(d_value, { batch }, opts);
        auto output = m_model->forward({ in }).toTuple();
        logits.copy_(output->elements()[0].toTensor());
        value.copy_(output->elements()[1].toTensor().reshape({ -1 }));
        return 0;
    }

    void setRowCount(const uint32_t* d_rows) override { m_dRows = d_rows; }

    // INetwork::evaluate for host callers (networks/GridNetwork.hpp:62-145 of the reference): the states are embedded as the
    // planes the search kernel writes (2t / 2t+1 = the mover's / the opponent's stones t plies back, last plane = player
    // ZERO to move, :72-97), the forward runs on the GPU, and exp / mask / normalise (uniform over the legal actions when
    // everything underflows, :107-142) follow on the host with GameActionDist's arithmetic.
    using ActionDist = GameActionDist<ACTION_SIZE>;
    std::vector<std::pair<ActionDist, Value>> evaluate(const std::vector<State>& states, const std::vector<ActionDist>& masks) override {
        std::vector<std::pair<ActionDist, Value>> results;
        const int64_t n = (int64_t)states.size();
        if (n == 0) return results;
        constexpr int CELLS = NUM_ROWS * NUM_COLS, PLANES = 2 * HISTORY_SIZE + 1;
        if (!m_model) throw EngineError(SPRL_E_STATE, "GridNetwork::evaluate: no model is loaded");
        if (m_planes == 0 && attach(currentDevice(), PLANES, NUM_ROWS, NUM_COLS, ACTION_SIZE) != 0)
            throw EngineError(SPRL_E_STATE, "GridNetwork::evaluate: the model could not be moved to the GPU");
        std::vector<float> planes((size_t)n * PLANES * CELLS, 0.0f);
        for (int64_t b = 0; b < n; ++b) {
            const State& st = states[(size_t)b];
            const Piece mine = pieceFromPlayer(st.getPlayer()), theirs = otherPiece(mine);
            float* p = planes.data() + (size_t)b * PLANES * CELLS;
            for (int t = 0; t < st.size() && t < HISTORY_SIZE; ++t)
                for (int c = 0; c < CELLS; ++c) {
                    p[(2 * t) * CELLS + c] = st.getHistory()[t][c] == mine ? 1.0f : 0.0f;
                    p[(2 * t + 1) * CELLS + c] = st.getHistory()[t][c] == theirs ? 1.0f : 0.0f;
                }
            if (st.getPlayer() == Player::ZERO) for (int c = 0; c < CELLS; ++c) p[(PLANES - 1) * CELLS + c] = 1.0f;
        }
        auto opts = torch::TensorOptions().dtype(torch::kFloat32).device(m_device);
        c10::cuda::CUDAStream stream = c10::cuda::getCurrentCUDAStream(m_device.index());
        c10::cuda::CUDAStreamGuard guard(stream);
        torch::Tensor in = torch::from_blob(planes.data(), { n, PLANES, NUM_ROWS, NUM_COLS }, torch::kFloat32).to(m_device);
        torch::Tensor logits = torch::empty({ n, ACTION_SIZE }, opts), value = torch::empty({ n }, opts);
        const uint32_t* saved = m_dRows;
        m_dRows = nullptr;                                   // the whole batch, not the engine's row count
        const int rc = forward(in.data_ptr<float>(), n, logits.data_ptr<float>(), value.data_ptr<float>(), stream.stream());
        m_dRows = saved;
        if (rc != 0) throw EngineError(rc, std::string("GridNetwork::evaluate: ") + sprl_last_error());
        const torch::Tensor hl = logits.cpu(), hv = value.cpu();
        check(checkStatus());
        results.reserve((size_t)n);
        for (int64_t b = 0; b < n; ++b) {
            ActionDist policy;
            for (int i = 0; i < ACTION_SIZE; ++i) policy[i] = hl.data_ptr<float>()[b * ACTION_SIZE + i];
            policy = policy.exp();
            int numLegal = 0;
            for (int i = 0; i < ACTION_SIZE; ++i) {
                if (masks[(size_t)b][i] == 0.0f) policy[i] = 0.0f; else ++numLegal;
            }
            const float sum = policy.sum();
            if (sum == 0.0f) {
                for (int i = 0; i < ACTION_SIZE; ++i) policy[i] = masks[(size_t)b][i] == 0.0f ? 0.0f : 1.0f / numLegal;
            } else {
                policy = policy / sum;
            }
            results.emplace_back(policy, hv.data_ptr<float>()[b]);
        }
        this->addEvals((uint64_t)n);
        return results;
    }
    int checkStatus() override { return m_evalnet ? sprl_evalnet_status(m_evalnet, nullptr) : 0; }

private:
    // Hands the module's parameters (names of src/networks/grid_networks.py:30-80) to the library's
    // evaluator; on any mismatch (board size, head shape) the TorchScript forward stays in use.
    void createEvalnet(int device, int planes, int rows, int cols, int actions) {
        std::map<std::string, torch::Tensor> sd;
        for (const auto& p : m_model->named_parameters()) sd[p.name] = p.value.detach().to(torch::kCPU).to(torch::kFloat32).contiguous();
        for (const auto& b : m_model->named_buffers()) sd[b.name] = b.value.detach().to(torch::kCPU).to(torch::kFloat32).contiguous();
        auto ptr = [&](const std::string& k) -> const float* { auto it = sd.find(k); return it == sd.end() ? nullptr : it->second.data_ptr<float>(); };
        auto convbn = [&](const std::string& c, const std::string& b) {
            sprl_conv_bn_params q { ptr(c + ".weight"), ptr(c + ".bias"), ptr(b + ".weight"), ptr(b + ".bias"), ptr(b + ".running_mean"), ptr(b + ".running_var") };
            return q;
        };
        if (!sd.count("conv.weight") || !sd.count("policy_fc.weight") || !sd.count("value_fc1.weight")) return;
        int blocks = 0;
        while (sd.count("residual_blocks." + std::to_string(blocks) + ".conv1.weight")) ++blocks;
        std::vector<sprl_conv_bn_params> tower;
        for (int i = 0; i < blocks; ++i) {
            const std::string b = "residual_blocks." + std::to_string(i);
            tower.push_back(convbn(b + ".conv1", b + ".bn1"));
            tower.push_back(convbn(b + ".conv2", b + ".bn2"));
        }
        sprl_network_params p {};
        p.rows = rows; p.cols = cols; p.in_planes = planes; p.channels = (int)sd["conv.weight"].size(0); p.blocks = blocks; p.actions = actions;
        p.policy_channels = (int)sd["policy_conv.weight"].size(0); p.value_channels = (int)sd["value_conv.weight"].size(0);
        p.value_hidden = (int)sd["value_fc1.weight"].size(0); p.bn_eps = 1e-5f;
        p.stem = convbn("conv", "bn"); p.tower = tower.data();
        p.policy_conv_w = ptr("policy_conv.weight"); p.policy_conv_b = ptr("policy_conv.bias");
        p.policy_fc_w = ptr("policy_fc.weight"); p.policy_fc_b = ptr("policy_fc.bias");
        p.value_conv_w = ptr("value_conv.weight"); p.value_conv_b = ptr("value_conv.bias");
        p.value_fc1_w = ptr("value_fc1.weight"); p.value_fc1_b = ptr("value_fc1.bias");
        p.value_fc2_w = ptr("value_fc2.weight"); p.value_fc2_b = ptr("value_fc2.bias");
        if (sprl_evalnet_create(device, &p, &m_evalnet) != SPRL_OK) {
            m_evalnet = nullptr;
            std::cout << "Evaluator: TorchScript module through LibTorch (" << sprl_last_error() << ")" << std::endl;
        } else {
            std::cout << "Evaluator: tcgen05 network kernel of libsprl_b200" << std::endl;
            // opt-in only: SPRL_EVALNET_PRECISION=fp16 selects the single-pass mode (not the reference's fp32 arithmetic)
            const char* prec = std::getenv("SPRL_EVALNET_PRECISION");
            if (prec && std::string(prec) == "fp16") {
                if (sprl_evalnet_set_precision(m_evalnet, SPRL_EVALNET_PRECISION_FP16) == SPRL_OK)
                    std::cout << "Evaluator precision: single-pass fp16 (SPRL_EVALNET_PRECISION=fp16; NOT the reference's fp32)" << std::endl;
                else
                    std::cout << "Evaluator precision: fp32-equivalent split (" << sprl_last_error() << ")" << std::endl;
            }
        }
    }

    sprl_evalnet* m_evalnet { nullptr };
    const uint32_t* m_dRows { nullptr };
    std::string m_path;
    torch::Device m_device { torch::kCPU };
    std::shared_ptr<torch::jit::script::Module> m_model;
    torch::Tensor m_input, m_logits, m_value;
    int m_planes { 0 }, m_rows { 0 }, m_cols { 0 }, m_actions { 0 };
};

}  // namespace SPRL
#endif
