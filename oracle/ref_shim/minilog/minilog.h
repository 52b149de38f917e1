/* Stand-in for corporateshark/minilog (un-vendored LVK dependency): log macros compile to nothing. Test infrastructure. */
#pragma once
#define LLOGL(...) ((void)0)
#define LLOGW(...) ((void)0)
#define LLOGD(...) ((void)0)
#define MINILOG_LOG_PROC(...) ((void)0)
