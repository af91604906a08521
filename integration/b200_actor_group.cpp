// B200ActorGroup — the binding a MiniZero maintainer adds to the reference tree to keep MiniZero's own ActorGroup (threads, wire protocol, actors,
// move decision, resignation, record writer) and hand only the per-move search to the B200 library through its C ABI (include/mz_b200.h).
//
// Seams used, all of them the reference's own virtuals (no reference source is modified):
//   ActorGroup::createNeuralNetworks  (actor/actor_group.h:55)   -> one mz_engine per GPU, weights through mz_net_*
//   ActorGroup::createActors          (actor/actor_group.h:56)   -> ZeroActor subclasses that can adopt a root table
//   BaseParalleler::newSlaveThread    (actor/actor_group.h:63)   -> slave threads whose CPU job is empty and whose GPU job is one whole move search
//   ActorGroup::handleCommand         (actor/actor_group.h:59)   -> load_model also reloads the engines
//   SlaveThread::handleSearchDone     (actor/actor_group.cpp:116-134), BaseActor::act, ThreadSharedData::outputGame: called unchanged
//
// Per move and GPU (slave thread id == GPU id, as in actor_group.cpp:99-114): draw the search's rotations and root noise with the thread's own
// utils::Random generator, run mz_search_run (all S + 1 cycles of actor_group.cpp:136-148 for the engine's games, one CUDA graph), read the root
// tables back, install each as the actor's MCTS root (MCTSNode setters, actor/mcts.h:29-41) and let the reference finish the move.
// PUCT searches (AlphaZero / board-game MuZero); Gumbel's move decision keeps state outside the tree (gumbel_zero.h:20-23) and stays with mz_sp.
//
// Built against the unmodified reference by oracle/Makefile (target b200) where /root/reference exists; tests/test_gpu_worker.py drives it over the
// wire protocol and lets the reference's own loader check the records.
#include "actor_group.h"
#include "configuration.h"
#include "configure_loader.h"
#include "create_network.h"
#include "environment.h"
#include "random.h"
#include "zero_actor.h"
#include <torch/cuda.h>
#include "../minizero_b200/host/net_loader.h" // the .pt reader of the drop-in worker (libtorch getters -> mz_net_*)
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

using namespace minizero;
using namespace minizero::actor;

#if GO
static const int kGame = MZ_GAME_GO;
#elif OTHELLO
static const int kGame = MZ_GAME_OTHELLO;
#elif NOGO
static const int kGame = MZ_GAME_NOGO;
#elif GOMOKU
static const int kGame = MZ_GAME_GOMOKU;
#elif HEX
static const int kGame = MZ_GAME_HEX;
#else
static const int kGame = MZ_GAME_TICTACTOE;
#endif

// A ZeroActor whose finished search comes from outside: the root and its children are written into the actor's own tree, then the reference's
// ZeroActor::handleSearchDone (zero_actor.cpp:160-176) chooses the move exactly as after its own search
class B200Actor : public ZeroActor {
public:
    using ZeroActor::ZeroActor;
    void adoptRoot(const mz_root_info& info, const int32_t* action, const float* count, const float* mean, const float* policy, const float* logit, const float* noise,
                   const float* value)
    {
        resetSearch(); // Tree::reset + the root's action (zero_actor.cpp:29-34)
        auto mcts = getMCTS();
        MCTSNode* root = mcts->getRootNode();
        root->setCount(info.count), root->setMean(info.mean), root->setValue(info.value);
        MCTSNode* child = mcts->allocateNodes(info.num_children);
        root->setFirstChild(child), root->setNumChildren(info.num_children);
        const env::Player turn = env_.getTurn();
        for (int i = 0; i < info.num_children; ++i) {
            child[i].reset();
            child[i].setAction(Action(action[i], turn));
            child[i].setCount(count[i]), child[i].setMean(mean[i]), child[i].setPolicy(policy[i]), child[i].setPolicyLogit(logit[i]);
            child[i].setPolicyNoise(noise[i]), child[i].setValue(value[i]);
        }
        handleSearchDone();
    }
    int numLegalActions() const { return static_cast<int>(env_.getLegalActions().size()); }
};

class B200SharedData : public ThreadSharedData {
public:
    std::vector<mz_engine*> engines_;
    std::vector<int> engine_games_;
};

class B200SlaveThread : public SlaveThread {
public:
    using SlaveThread::SlaveThread;

protected:
    bool doCPUJob() override { return false; } // selection, environment, expansion and backup all run on the device
    void doGPUJob() override
    {
        auto sd = std::static_pointer_cast<B200SharedData>(shared_data_);
        const int ne = static_cast<int>(sd->engines_.size());
        if (id_ >= ne) { return; }
        mz_engine* eng = sd->engines_[id_];
        const int n = sd->engine_games_[id_], A = mz_action_size(eng), S1 = config::actor_num_simulation + 1;
        const bool muzero = (config::nn_type_name == "muzero");
        // randomness of one search, from this thread's generator (actor_group.cpp:66-70): root noise per game, one rotation per game and cycle
        std::vector<uint8_t> rot(static_cast<size_t>(S1) * n, 0);
        std::vector<float> noise(static_cast<size_t>(n) * A, 0.0f);
        for (int slot = 0; slot < n; ++slot) {
            auto actor = std::static_pointer_cast<B200Actor>(sd->actors_[slot * ne + id_]);
            if (config::actor_use_dirichlet_noise) {
                const std::vector<float> dir = utils::Random::randDirichlet(config::actor_dirichlet_noise_alpha, actor->numLegalActions());
                std::copy(dir.begin(), dir.end(), noise.begin() + static_cast<size_t>(slot) * A);
            }
            if (config::actor_use_random_rotation_features && !muzero) {
                for (int c = 0; c < S1; ++c) { rot[static_cast<size_t>(c) * n + slot] = static_cast<uint8_t>(utils::Random::randInt() % 8); }
            }
        }
        check(mz_search_set_inputs(eng, (config::actor_use_random_rotation_features && !muzero) ? rot.data() : nullptr, config::actor_use_dirichlet_noise ? noise.data() : nullptr));
        check(mz_search_run(eng, 0, nullptr));
        std::vector<mz_root_info> info(n);
        std::vector<int32_t> action(static_cast<size_t>(n) * A);
        std::vector<float> f[6];
        for (auto& v : f) { v.resize(static_cast<size_t>(n) * A); }
        check(mz_get_roots(eng, info.data(), action.data(), f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data()));
        // the reference finishes every move: decision, resignation, act, records, next game (actor_group.cpp:116-134)
        std::vector<int32_t> play(n, -1);
        std::vector<int> restarted;
        for (int slot = 0; slot < n; ++slot) {
            const int actor_id = slot * ne + id_;
            auto actor = std::static_pointer_cast<B200Actor>(sd->actors_[actor_id]);
            const size_t o = static_cast<size_t>(slot) * A;
            actor->adoptRoot(info[slot], action.data() + o, f[0].data() + o, f[1].data() + o, f[2].data() + o, f[3].data() + o, f[4].data() + o, f[5].data() + o);
            const bool resign = actor->isResign();
            const int chosen = actor->getSearchAction().getActionID();
            handleSearchDone(actor_id);
            if (!resign) { play[slot] = chosen; }
            if (actor->getEnvironment().getActionHistory().empty()) { restarted.push_back(slot); } // the game ended: the actor was reset
        }
        std::vector<mz_play_result> res(n);
        check(mz_play(eng, play.data(), res.data()));
        for (int slot = 0; slot < n; ++slot) {
            if (play[slot] >= 0 && !res[slot].applied) {
                std::cerr << "device rejected action " << play[slot] << std::endl;
                exit(-1);
            }
        }
        for (int slot : restarted) { check(mz_reset_game(eng, slot)); }
    }

private:
    static void check(int rc)
    {
        if (rc != MZ_OK) {
            std::cerr << "libmzb200: " << mz_last_error() << std::endl;
            exit(-1);
        }
    }
};

class B200ActorGroup : public ActorGroup {
protected:
    void createSharedData() override { shared_data_ = std::make_shared<B200SharedData>(); }
    std::shared_ptr<utils::BaseSlaveThread> newSlaveThread(int id) override { return std::make_shared<B200SlaveThread>(id, shared_data_); }
    std::shared_ptr<B200SharedData> b200() { return std::static_pointer_cast<B200SharedData>(shared_data_); }

    void createNeuralNetworks() override
    {
        // the reference's Network object stays (on the CPU) for what the actors ask it: type name, action size (zero_actor.cpp:100-114)
        getSharedData()->networks_.resize(1);
        getSharedData()->network_outputs_.resize(1);
        getSharedData()->networks_[0] = network::createNetwork(config::nn_file_name, -1);
        const int ne = std::min(static_cast<int>(torch::cuda::device_count()), config::zero_num_parallel_games);
        if (ne < 1) {
            std::cerr << "no CUDA device" << std::endl;
            exit(-1);
        }
        for (int e = 0; e < ne; ++e) {
            const int games = config::zero_num_parallel_games / ne + (e < config::zero_num_parallel_games % ne ? 1 : 0); // actor i -> engine i % ne (actor_group.cpp:184-186)
            mz_config c{};
            c.device = e, c.game = kGame, c.board_size = config::env_board_size, c.num_games = games, c.num_simulation = config::actor_num_simulation;
            c.puct_base = config::actor_mcts_puct_base, c.puct_init = config::actor_mcts_puct_init, c.reward_discount = config::actor_mcts_reward_discount;
            c.dirichlet_epsilon = config::actor_dirichlet_noise_epsilon, c.muzero = (config::nn_type_name == "muzero");
            c.gomoku_exactly_five = 1, c.hex_swap_rule = 1; // reference defaults (configuration.cpp:82-85); the per-game keys exist only in their builds
#if GO || NOGO
            c.komi = config::env_go_komi, c.ko_situational = (config::env_go_ko_rule == "situational");
#endif
            mz_engine* eng = nullptr;
            if (mz_create(&c, &eng) != MZ_OK) {
                std::cerr << "mz_create: " << mz_last_error() << std::endl;
                exit(-1);
            }
            b200()->engines_.push_back(eng);
            b200()->engine_games_.push_back(games);
        }
        loadEngines();
    }

    void loadEngines()
    {
        for (mz_engine* eng : b200()->engines_) {
            std::string error;
            if (!mzhost::loadNetwork(config::nn_file_name, eng, error)) {
                std::cerr << "load_model: " << error << std::endl;
                exit(-1);
            }
        }
    }

    void createActors() override
    {
        std::shared_ptr<network::Network>& network = getSharedData()->networks_[0];
        const uint64_t tree_node_size = static_cast<uint64_t>(network->getActionSize()) + 1; // the actor's tree only ever holds a root and its children
        for (int i = 0; i < config::zero_num_parallel_games; ++i) {
            auto actor = std::make_shared<B200Actor>(tree_node_size);
            actor->setNetwork(network);
            actor->reset();
            getSharedData()->actors_.emplace_back(actor);
        }
    }

    void handleCommand(const std::string& command_prefix, const std::string& command) override
    {
        ActorGroup::handleCommand(command_prefix, command);
        if (command_prefix == "load_model") { loadEngines(); }
        if (command_prefix == "reset_actors") {
            for (mz_engine* eng : b200()->engines_) { mz_reset_game(eng, -1); }
        }
    }
};

int main(int argc, char** argv)
{
    if (argc < 2) {
        std::cerr << "usage: b200_actor_group <conf_str>     (speaks the zero-server wire protocol on stdin / stdout, like `-mode sp`)" << std::endl;
        return 2;
    }
    env::setUpEnv();
    config::ConfigureLoader cl;
    config::setConfiguration(cl);
    if (!cl.loadFromString(argv[1])) { return 1; }
    utils::Random::seed(config::program_seed);
    B200ActorGroup ag;
    ag.run();
    return 0;
}
