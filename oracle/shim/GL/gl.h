/* Shim: the CUDA toolkit's cuda_gl_interop.h includes <GL/gl.h>, which this image does not ship.
 * The reference vendors glad (include/glad/glad.h), which supplies the GL types it needs. */
#include <glad/glad.h>
