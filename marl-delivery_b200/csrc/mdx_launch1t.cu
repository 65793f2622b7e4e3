// stale-tracker variants (MDX_CFG_STALE_TRACKER) of kernel mode 1
#define MDX_MODE 1
#define MDX_TRK 1
#include "mdx_launch.inl"
