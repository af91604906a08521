// Configuration of the self-play worker: the reference's `key=value` registry (config/configuration.cpp:92-205,
// config/configure_loader.cpp:34-108). Every key the reference registers is accepted (an unknown key makes loading
// fail, as there); only the keys the actor path reads are interpreted.
#pragma once
#include <map>
#include <string>

namespace mzhost {

class Config {
public:
    Config();
    bool loadFromFile(const std::string& path);  // configure_loader.cpp:34-48
    bool loadFromString(const std::string& s);   // configure_loader.cpp:50-61 (':'-separated)
    bool setValue(std::string line);             // configure_loader.cpp:90-114
    int getInt(const std::string& k) const;
    float getFloat(const std::string& k) const;
    bool getBool(const std::string& k) const;
    const std::string& getString(const std::string& k) const;
    void set(const std::string& k, const std::string& v) { values_[k] = v; }

private:
    std::map<std::string, std::string> values_; // key -> textual value (defaults of configuration.cpp:6-90)
    std::map<std::string, char> types_;         // 'i' int, 'f' float, 'b' bool, 's' string
};

} // namespace mzhost
