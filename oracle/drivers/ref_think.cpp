// Golden-vector generator for the console search: drives the UNMODIFIED reference ZeroActor::think()
// (actor/zero_actor.cpp:36-49,129-157) with actor_mcts_think_batch_size = K and records, through the actor's own virtual
// hooks, every selection of every batched step (rotation, path length, whether the leaf joins the batch, feature
// planes) and every network output the actor consumes, plus the root child table each search ends with.
// TEST INFRASTRUCTURE ONLY.
//
// usage: ref_think <conf_str> <out_dir> <moves>
#include "configuration.h"
#include "configure_loader.h"
#include "create_network.h"
#include "environment.h"
#include "random.h"
#include "zero_actor.h"
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

using namespace minizero;
using namespace minizero::actor;
using namespace minizero::network;

static FILE* f_ev = nullptr;
static int g_A = 0, g_F = 0;
static void put_i32(FILE* f, int32_t v) { fwrite(&v, 4, 1, f); }
static void put_f32(FILE* f, float v) { fwrite(&v, 4, 1, f); }

class ProbeActor : public ZeroActor {
public:
    using ZeroActor::ZeroActor;
    void beforeNNEvaluation() override
    {
        ZeroActor::beforeNNEvaluation();
        const auto& path = mcts_search_data_.node_path_;
        // record kind 0: batch id, rotation, path length, virtual loss of the leaf BEFORE this selection's own is added (0: it joins the batch)
        put_i32(f_ev, 0), put_i32(f_ev, nn_evaluation_batch_id_), put_i32(f_ev, alphazero_network_ ? static_cast<int>(feature_rotation_) : 0), put_i32(f_ev, static_cast<int>(path.size()));
        put_f32(f_ev, path.back()->getVirtualLoss());
        std::vector<uint8_t> fb(g_F, 0);
        if (alphazero_network_) {
            Environment t = getEnvironmentTransition(path);
            std::vector<float> feats = t.getFeatures(feature_rotation_);
            for (int k = 0; k < g_F; ++k) { fb[k] = (feats[k] != 0.0f); }
        } else if (path.size() == 1) { // MuZero: only the root's initial inference consumes planes (zero_actor.cpp:59-61)
            std::vector<float> feats = env_.getFeatures();
            for (int k = 0; k < g_F; ++k) { fb[k] = (feats[k] != 0.0f); }
        }
        fwrite(fb.data(), 1, g_F, f_ev);
    }
    void afterNNEvaluation(const std::shared_ptr<NetworkOutput>& out) override
    {
        // record kind 1: batch id, policy, logits, value
        put_i32(f_ev, 1), put_i32(f_ev, nn_evaluation_batch_id_);
        if (alphazero_network_) {
            auto o = std::static_pointer_cast<AlphaZeroNetworkOutput>(out);
            fwrite(o->policy_.data(), 4, g_A, f_ev);
            fwrite(o->policy_logits_.data(), 4, g_A, f_ev);
            put_f32(f_ev, o->value_);
        } else {
            auto o = std::static_pointer_cast<MuZeroNetworkOutput>(out);
            fwrite(o->policy_.data(), 4, g_A, f_ev);
            fwrite(o->policy_logits_.data(), 4, g_A, f_ev);
            put_f32(f_ev, o->value_);
        }
        ZeroActor::afterNNEvaluation(out);
    }
};

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::cerr << "usage: ref_think <conf_str> <out_dir> <moves>" << std::endl;
        return 2;
    }
    const std::string out_dir = argv[2];
    const int moves = atoi(argv[3]);
    env::setUpEnv();
    config::ConfigureLoader cl;
    config::setConfiguration(cl);
    if (!cl.loadFromString(argv[1])) { return 1; }
    utils::Random::seed(config::program_seed);
    std::shared_ptr<Network> network = createNetwork(config::nn_file_name, -1);
    if (network->getNetworkTypeName() != "alphazero" && network->getNetworkTypeName() != "muzero") { return 3; }
    g_A = network->getActionSize();
    g_F = network->getNumInputChannels() * network->getInputChannelHeight() * network->getInputChannelWidth();
    const uint64_t tree_node_size = static_cast<uint64_t>(config::actor_num_simulation + 1) * g_A;
    auto actor = std::make_shared<ProbeActor>(tree_node_size);
    actor->setNetwork(network);
    actor->reset();

    f_ev = fopen((out_dir + "/events.bin").c_str(), "wb");
    FILE* f_move = fopen((out_dir + "/moves.bin").c_str(), "wb");
    FILE* f_meta = fopen((out_dir + "/meta.txt").c_str(), "w");
    fprintf(f_meta, "A %d\nF %d\nS %d\nK %d\n", g_A, g_F, config::actor_num_simulation, config::actor_mcts_think_batch_size);
    fclose(f_meta);
    for (int m = 0; m < moves && !actor->isEnvTerminal(); ++m) {
        const Action action = actor->think(false, false);
        put_i32(f_ev, 2); // record kind 2: the search is over
        const MCTSNode* root = actor->getMCTS()->getRootNode();
        put_i32(f_move, action.getActionID());
        put_i32(f_move, static_cast<int32_t>(action.getPlayer()));
        put_i32(f_move, root->getNumChildren());
        put_f32(f_move, root->getCount());
        put_f32(f_move, root->getMean());
        put_f32(f_move, root->getValue());
        for (int c = 0; c < g_A; ++c) {
            const MCTSNode* ch = (c < root->getNumChildren() ? root->getChild(c) : nullptr);
            put_i32(f_move, ch ? ch->getAction().getActionID() : -1);
            put_f32(f_move, ch ? ch->getCount() : 0.f);
            put_f32(f_move, ch ? ch->getMean() : 0.f);
            put_f32(f_move, ch ? ch->getPolicy() : 0.f);
            put_f32(f_move, ch ? ch->getPolicyLogit() : 0.f);
            put_f32(f_move, ch ? ch->getPolicyNoise() : 0.f);
            put_f32(f_move, ch ? ch->getValue() : 0.f);
            put_f32(f_move, ch ? ch->getVirtualLoss() : 0.f);
        }
        actor->act(action);
    }
    fclose(f_ev);
    fclose(f_move);
    return 0;
}
