from .. import extend_with_reference

extend_with_reference(__path__, "utils")
