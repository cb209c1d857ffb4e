/* Return codes of the C binding (ABI of the reference's binding/c/include/AC/Error.h:4-9). */
#ifndef AC_BINDING_C_ERROR_H
#define AC_BINDING_C_ERROR_H

#define AC_SUCCESS 0
#define AC_ERROR(e) (-(e))

#define AC_EIO          5     /* file I/O failed */
#define AC_EINVAL      22     /* NULL or unusable argument */
#define AC_EPROCESSOR 256     /* the processor reported !ok() */

#endif
