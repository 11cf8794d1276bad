// empty: the reference includes this for the CV_* constants, which the opencv.hpp stub already defines.
