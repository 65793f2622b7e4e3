#define MDX_MODE 1
#include "mdx_launch.inl"
