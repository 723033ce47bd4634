// stand-in include: see standin_ros_core.hpp
#include "standin_ros_core.hpp"
