#define MDX_MODE 2
#include "mdx_launch.inl"
