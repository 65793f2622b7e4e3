// stale-tracker variants (MDX_CFG_STALE_TRACKER) of kernel mode 2
#define MDX_MODE 2
#define MDX_TRK 1
#include "mdx_launch.inl"
