// Stand-in for the autoconf-generated config.h (configure.ac:2 -> 2.1.2).
#define PACKAGE_VERSION "2.1.2"
#define SVN_REVISION "oracle"
