// Drop-in check with the reference's own code: reads `SelfPlay ...` lines (as the zero server receives them) from stdin,
// parses every record with the reference's EnvironmentLoader (environment/base/base_env.h:149-205), replays it through
// the reference's Environment (act must accept every move), and compares the terminal flag / RE tag / return / lengths
// with what the reference itself computes. Prints "RECORDS_OK <n>" or the first problem. TEST INFRASTRUCTURE ONLY.
//
// usage: ref_record_check <conf_str>      (e.g. "env_board_size=9")
#include "configuration.h"
#include "configure_loader.h"
#include "environment.h"
#include <iostream>
#include <sstream>
#include <string>

using namespace minizero;

int main(int argc, char** argv)
{
    env::setUpEnv();
    config::ConfigureLoader cl;
    config::setConfiguration(cl);
    if (argc > 1 && std::string(argv[1]).size() && !cl.loadFromString(argv[1])) { return 2; }
    std::string line;
    int n = 0;
    while (std::getline(std::cin, line)) {
        if (line.empty()) { continue; }
        // zero_server.cpp:39-52,111-114: "SelfPlay <terminal> <data_len> <game_len> <return> <record> #"
        if (line.rfind("SelfPlay ", 0) != 0 || line.size() < 2 || line.substr(line.size() - 2) != " #") {
            std::cout << "BAD_LINE " << n << std::endl;
            return 1;
        }
        std::istringstream iss(line);
        std::string tag, is_terminal, record;
        int data_len, game_len;
        float ret;
        iss >> tag >> is_terminal >> data_len >> game_len >> ret >> record;
        EnvironmentLoader loader;
        if (!loader.loadFromString(record)) {
            std::cout << "PARSE_FAIL " << n << std::endl;
            return 1;
        }
        Environment e;
        e.reset();
        for (auto& p : loader.getActionPairs()) {
            if (!e.act(p.first)) {
                std::cout << "ILLEGAL_MOVE " << n << " action " << p.first.getActionID() << std::endl;
                return 1;
            }
            if (p.second.count("P") == 0 || p.second.count("V") == 0 || p.second.count("R") == 0) {
                std::cout << "MISSING_INFO " << n << std::endl;
                return 1;
            }
        }
        const int len = static_cast<int>(loader.getActionPairs().size());
        if (len != game_len || data_len != game_len) {
            std::cout << "LENGTH_MISMATCH " << n << std::endl;
            return 1;
        }
        const float expect = e.getEvalScore(!e.isTerminal());
        if (expect != ret || std::stof(loader.getTag("RE")) != expect) {
            std::cout << "RESULT_MISMATCH " << n << " expected " << expect << " got " << ret << " RE " << loader.getTag("RE") << std::endl;
            return 1;
        }
        ++n;
    }
    std::cout << "RECORDS_OK " << n << std::endl;
    return 0;
}
