// stub: the hot-path translation units use nothing from this header
#pragma once
