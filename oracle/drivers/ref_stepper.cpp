// Golden-vector generator: drives the UNMODIFIED reference actor/network/environment code
// (compiled in place from /root/reference by oracle/Makefile) through a deterministic
// single-thread restatement of the ActorGroup cycle (actor/actor_group.cpp:81-134,136-148) and
// dumps, for every leaf evaluation, the feature planes pushed to the network and the network's
// outputs, and for every move the root child table. TEST INFRASTRUCTURE ONLY.
//
// usage: ref_stepper <conf_str> <out_dir> <max_moves_total> [max_cycles]
#include "actor_group.h"
#include "configuration.h"
#include "configure_loader.h"
#include "create_network.h"
#include "environment.h"
#include "random.h"
#include "zero_actor.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

using namespace minizero;
using namespace minizero::actor;
using namespace minizero::network;

class ProbeActor : public ZeroActor {
public:
    using ZeroActor::ZeroActor;
    const std::vector<MCTSNode*>& nodePath() const { return mcts_search_data_.node_path_; }
    int rotation() const { return static_cast<int>(feature_rotation_); }
    Environment transition() { return getEnvironmentTransition(mcts_search_data_.node_path_); }
    bool resignEnabled() const { return enable_resign_; }
};

static void put_i32(FILE* f, int32_t v) { fwrite(&v, 4, 1, f); }
static void put_f32(FILE* f, float v) { fwrite(&v, 4, 1, f); }

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::cerr << "usage: ref_stepper <conf_str> <out_dir> <max_moves_total> [max_cycles]" << std::endl;
        return 2;
    }
    const std::string out_dir = argv[2];
    const long max_moves = atol(argv[3]);
    const long max_cycles = (argc > 4 ? atol(argv[4]) : -1);

    env::setUpEnv();
    config::ConfigureLoader cl;
    config::setConfiguration(cl);
    if (!cl.loadFromString(argv[1])) { return 1; }
    utils::Random::seed(config::program_seed);

    std::shared_ptr<Network> network = createNetwork(config::nn_file_name, -1);
    const bool is_az = (network->getNetworkTypeName() == "alphazero");
    const bool is_atari = (network->getNetworkTypeName() == "muzero_atari"); // discrete value / reward heads, value rescale, rewards in the tree
    const int A = network->getActionSize();
    const int F = network->getNumInputChannels() * network->getInputChannelHeight() * network->getInputChannelWidth();
    const uint64_t tree_node_size = static_cast<uint64_t>(config::actor_num_simulation + 1) * A;

    std::vector<std::shared_ptr<ProbeActor>> actors;
    for (int i = 0; i < config::zero_num_parallel_games; ++i) {
        auto a = std::make_shared<ProbeActor>(tree_node_size);
        a->setNetwork(network);
        a->reset(); // one randReal() per actor on the main-thread generator (create_actor.h:12-14)
        actors.push_back(a);
    }
    // slave thread 0 re-seeds its own generator with program_seed + 0 (actor_group.cpp:66-70)
    utils::Random::seed(config::program_seed);

    FILE* f_eval = fopen((out_dir + "/evals.bin").c_str(), "wb");
    FILE* f_move = fopen((out_dir + "/moves.bin").c_str(), "wb");
    FILE* f_root = (is_atari ? fopen((out_dir + "/roots.bin").c_str(), "wb") : nullptr);
    FILE* f_meta = fopen((out_dir + "/meta.txt").c_str(), "w");
    fprintf(f_meta, "A %d\nF %d\nS %d\nB %d\ntype %s\n", A, F, config::actor_num_simulation, config::zero_num_parallel_games, network->getNetworkTypeName().c_str());
    fclose(f_meta);

    ThreadSharedData shared; // only for outputGame(): writes the SelfPlay line to stdout
    std::vector<std::shared_ptr<NetworkOutput>> outputs;
    std::vector<std::vector<float>> cycle_feats(actors.size());
    std::vector<std::vector<int32_t>> cycle_hdr(actors.size());
    long moves = 0;
    for (long cycle = 0; max_cycles < 0 || cycle < max_cycles; ++cycle) {
        bool stop = false;
        for (size_t i = 0; i < actors.size(); ++i) {
            auto& actor = actors[i];
            int idx = actor->getNNEvaluationBatchIndex();
            if (idx >= 0) {
                actor->afterNNEvaluation(outputs[idx]);
                if (actor->isSearchDone()) {
                    // root child table at the moment the move is decided
                    auto mcts = actor->getMCTS();
                    const MCTSNode* root = mcts->getRootNode();
                    const bool resign = actor->isResign();
                    const Action action = actor->getSearchAction();
                    put_i32(f_move, static_cast<int32_t>(i));
                    put_i32(f_move, static_cast<int32_t>(actor->getEnvironment().getActionHistory().size()));
                    put_i32(f_move, action.getActionID());
                    put_i32(f_move, static_cast<int32_t>(action.getPlayer()));
                    put_i32(f_move, root->getNumChildren());
                    put_i32(f_move, resign ? 1 : 0);
                    put_f32(f_move, root->getCount());
                    put_f32(f_move, root->getMean());
                    put_f32(f_move, root->getValue());
                    if (is_atari) { // the value bounds the search ended with (mcts.h:106, used by the move choice and the resign test)
                        const auto& vb = mcts->getTreeValueBound();
                        put_i32(f_move, static_cast<int32_t>(vb.size()));
                        put_f32(f_move, vb.empty() ? 0.f : vb.begin()->first);
                        put_f32(f_move, vb.empty() ? 0.f : vb.rbegin()->first);
                    }
                    for (int c = 0; c < A; ++c) {
                        const MCTSNode* ch = (c < root->getNumChildren() ? root->getChild(c) : nullptr);
                        put_i32(f_move, ch ? ch->getAction().getActionID() : -1);
                        put_f32(f_move, ch ? ch->getCount() : 0.f);
                        put_f32(f_move, ch ? ch->getMean() : 0.f);
                        put_f32(f_move, ch ? ch->getPolicy() : 0.f);
                        put_f32(f_move, ch ? ch->getPolicyLogit() : 0.f);
                        put_f32(f_move, ch ? ch->getPolicyNoise() : 0.f);
                        put_f32(f_move, ch ? ch->getValue() : 0.f);
                        if (is_atari) { put_f32(f_move, ch ? ch->getReward() : 0.f); }
                    }
                    // SlaveThread::handleSearchDone (actor_group.cpp:116-134)
                    if (!resign) { actor->act(action); }
                    if (is_atari) { // what the environment answered: reward of the move, terminal flag (the frame itself is in the next root's planes)
                        put_f32(f_move, actor->getEnvironment().getReward());
                        put_f32(f_move, actor->getEnvironment().getEvalScore());
                        put_i32(f_move, actor->isEnvTerminal() ? 1 : 0);
#if ATARI
                        put_i32(f_move, actor->getEnvironment().getSeed());
                        put_i32(f_move, actor->getEnvironment().getLives());
#else
                        put_i32(f_move, 0), put_i32(f_move, 0);
#endif
                    }
                    bool is_endgame = (resign || actor->isEnvTerminal());
                    if (is_endgame) {
                        shared.outputGame(actor);
                        actor->reset();
                    } else {
                        int game_length = actor->getEnvironment().getActionHistory().size();
                        int sequence_length = config::zero_actor_intermediate_sequence_length;
                        if (sequence_length > 0 && game_length >= sequence_length && (game_length - config::learner_n_step_return - config::learner_muzero_unrolling_step) % sequence_length == 0) { shared.outputGame(actor); }
                        actor->resetSearch();
                    }
                    if (++moves >= max_moves) { stop = true; }
                }
            }
            if (stop) { break; }
            actor->beforeNNEvaluation();
            // what was pushed: recompute the same features the actor just handed to the network
            const auto& path = actor->nodePath();
            cycle_hdr[i] = {static_cast<int32_t>(cycle), static_cast<int32_t>(i), is_az ? actor->rotation() : 0, static_cast<int32_t>(path.size())};
            if (!is_az) {
                // MuZero: the leaf is identified by its action and the path by a hash of its action ids (records of the
                // muzero type carry a 6-int header; alphazero records keep the 4-int header of the committed fixtures)
                uint32_t h = 2166136261u;
                for (size_t k = 1; k < path.size(); ++k) { h = (h ^ static_cast<uint32_t>(path[k]->getAction().getActionID())) * 16777619u; }
                cycle_hdr[i].push_back(path.back()->getAction().getActionID());
                cycle_hdr[i].push_back(static_cast<int32_t>(h & 0x7fffffffu));
            }
            if (is_az) {
                Environment t = actor->transition();
                cycle_feats[i] = t.getFeatures(static_cast<utils::Rotation>(actor->rotation()));
            } else {
                cycle_feats[i].assign(F, 0.0f);
                if (path.size() == 1) { cycle_feats[i] = actor->getEnvironment().getFeatures(); }
            }
        }
        if (stop) { break; }
        if (is_az) {
            outputs = std::static_pointer_cast<AlphaZeroNetwork>(network)->forward();
        } else {
            auto mz = std::static_pointer_cast<MuZeroNetwork>(network);
            outputs = (mz->getInitialInputBatchSize() > 0 ? mz->initialInference() : mz->recurrentInference());
        }
        for (size_t i = 0; i < actors.size(); ++i) {
            int idx = actors[i]->getNNEvaluationBatchIndex();
            fwrite(cycle_hdr[i].data(), 4, cycle_hdr[i].size(), f_eval);
            std::vector<uint8_t> fb(F);
            if (is_atari) {
                // planes are bytes / 255 (RGB) or action id / 18 (atari.cpp:82,156): recorded as those integers, the test rebuilds the floats.
                // Only root evaluations push planes (zero_actor.cpp:59-61); they go to roots.bin as (cycle, game, F bytes)
                if (cycle_hdr[i][3] == 1) {
                    const int hw = network->getInputChannelHeight() * network->getInputChannelWidth();
                    for (int k = 0; k < F; ++k) { fb[k] = static_cast<uint8_t>(std::lround(cycle_feats[i][k] * (((k / hw) % 4 == 0) ? 18.0f : 255.0f))); }
                    fwrite(cycle_hdr[i].data(), 4, 2, f_root);
                    fwrite(fb.data(), 1, F, f_root);
                }
            } else {
                for (int k = 0; k < F; ++k) { fb[k] = (cycle_feats[i][k] != 0.0f); }
                fwrite(fb.data(), 1, F, f_eval);
            }
            if (is_az) {
                auto o = std::static_pointer_cast<AlphaZeroNetworkOutput>(outputs[idx]);
                fwrite(o->policy_.data(), 4, A, f_eval);
                fwrite(o->policy_logits_.data(), 4, A, f_eval);
                put_f32(f_eval, o->value_);
            } else {
                auto o = std::static_pointer_cast<MuZeroNetworkOutput>(outputs[idx]);
                fwrite(o->policy_.data(), 4, A, f_eval);
                fwrite(o->policy_logits_.data(), 4, A, f_eval);
                put_f32(f_eval, o->value_);
                if (is_atari) { put_f32(f_eval, o->reward_); } // after the expectation over the 601 bins and invertValue (muzero_network.h:157-171)
            }
        }
    }
    fclose(f_eval);
    fclose(f_move);
    if (f_root) { fclose(f_root); }
    return 0;
}
