from .filtering import moment_filter_rms, moment_filter_cms, moment_filter_scms
from .quadtures import moment_quadrature
