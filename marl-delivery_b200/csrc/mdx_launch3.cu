#define MDX_MODE 3
#include "mdx_launch.inl"
