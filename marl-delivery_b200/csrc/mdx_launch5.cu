#define MDX_MODE 5
#include "mdx_launch.inl"
