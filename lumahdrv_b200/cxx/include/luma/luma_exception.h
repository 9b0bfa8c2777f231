/*
 * luma_exception.h -- error type of the drop-in C++ facade.
 *
 * Same name, constructor and what() as the reference's LumaException
 * (reference include/luma/luma_exception.h:53-71), so that the reference's
 * drivers (lumaenc.cpp:247-262, lumadec.cpp) catch it unchanged.  CUDA-side
 * failures reported by the C ABI (include/lumacu.h) are rethrown as this type.
 */
#ifndef LUMA_EXCEPTION_H
#define LUMA_EXCEPTION_H

#include <stdexcept>
#include <string>

class LumaException : public std::exception
{
public:
    LumaException(const char *message) : m_text(message ? message : "") {}
    explicit LumaException(const std::string &message) : m_text(message) {}
    virtual ~LumaException() throw() {}
    virtual const char *what() const throw() { return m_text.c_str(); }

private:
    std::string m_text;
};

#endif // LUMA_EXCEPTION_H
