// oracle shim: forwards to the single cv::Mat stand-in (test infrastructure only).
#include <opencv2/core/core.hpp>
