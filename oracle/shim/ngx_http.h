/* TEST INFRASTRUCTURE — intentionally empty stand-in (see ngx_config.h). */
