#pragma once
#include "../filtering_streambuf.hpp"
