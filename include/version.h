#ifndef HPRLP_VERSION_H
#define HPRLP_VERSION_H
/* Version macros kept equal to the reference include/version.h:13-21 (API level 0.1.2). */
#define HPRLP_VERSION_MAJOR 0
#define HPRLP_VERSION_MINOR 1
#define HPRLP_VERSION_PATCH 0
#define HPRLP_VERSION_STRING "0.1.2"
#define HPRLP_VERSION_NUMBER 101
#define HPRLP_ENGINE_STRING "b200-native r1"
#endif
