// STUB (oracle/stub): see pcl/point_types.h
#pragma once
#include <pcl/point_types.h>
