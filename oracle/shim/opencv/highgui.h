/* TEST INFRASTRUCTURE — intentionally empty stand-in; codec prototypes live in cv.h. */
